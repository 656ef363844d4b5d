"""The reference's own testsets for the hot path (test/schemes.jl, test/models.jl), run through
the product path on the GPU with the same model, chi, iteration count and tolerance."""
import numpy as np
import pytest

import tnr_oracle as o

pytestmark = pytest.mark.gpu


def rel(a, b):
    return abs((a - b) / b)


@pytest.mark.parametrize("name,chi,sf,tol", [
    ("TRG", 24, 2.0, 2.0e-6),     # test/schemes.jl:19-26
    ("BTRG", 24, 2.0, 6.0e-8),    # test/schemes.jl:61-66
    ("HOTRG", 16, 4.0, 6.0e-7),   # test/schemes.jl:99-104
    ("ATRG", 24, 4.0, 3.0e-6),    # test/schemes.jl:137-142
])
def test_2d_free_energy_testsets(tk, name, chi, sf, tol):
    T = tk.classical_ising()  # Z2-symmetric model, as in the reference tests
    data = tk.run(getattr(tk, name)(T), tk.truncrank(chi), tk.maxiter(25), verbosity=0)
    assert len(data) == 26
    assert rel(tk.free_energy(data, tk.ising_βc, scalefactor=sf), tk.f_onsager) < tol


def test_hotrg3d_testset(tk):
    # test/schemes.jl:366-373: HOTRG_3D, truncrank(8), maxiter(25), rtol 1e-3 vs -3.507
    data = tk.run(tk.HOTRG_3D(tk.classical_ising_3D()), tk.truncrank(8), tk.maxiter(25), verbosity=0)
    assert rel(tk.free_energy(data, tk.ising_βc_3D, scalefactor=8.0), -3.507) < 1.0e-3


def test_atrg3d_testset(tk):
    # test/schemes.jl:356-363: ATRG_3D, truncrank(12), maxiter(25), rtol 5e-3 vs -3.507
    data = tk.run(tk.ATRG_3D(tk.classical_ising_3D()), tk.truncrank(12), tk.maxiter(25), verbosity=0)
    assert rel(tk.free_energy(data, tk.ising_βc_3D, scalefactor=8.0), -3.507) < 5.0e-3


@pytest.mark.parametrize("model,beta,answer", [
    ("ising_trivial", None, None), ("ising_z2", None, None),
    ("potts_trivial", None, -4.119552029995684), ("potts_z3", None, -4.119552029995684),
])
def test_models_testset(tk, model, beta, answer):
    # test/models.jl:30-36: TRG, truncrank(16), maxiter(25), rtol 1e-3
    T = {"ising_trivial": lambda: tk.classical_ising(tk.Trivial),
         "ising_z2": lambda: tk.classical_ising(),
         "potts_trivial": lambda: tk.classical_potts(tk.Trivial, 3),
         "potts_z3": lambda: tk.classical_potts(3)}[model]()
    beta = tk.ising_βc if model.startswith("ising") else tk.potts_βc(3)
    answer = tk.f_onsager if answer is None else answer
    data = tk.run(tk.TRG(T), tk.truncrank(16), tk.maxiter(25), verbosity=0)
    assert rel(tk.free_energy(data, beta), answer) < 1.0e-3


def test_3d_models_testset(tk):
    # test/models.jl:25-28,80-86: HOTRG_3D, truncrank(8), maxiter(25), rtol 1e-3 vs -3.508 for the
    # Trivial and the Z2 tensor
    for T in (tk.classical_ising_3D(tk.Trivial), tk.classical_ising_3D()):
        data = tk.run(tk.HOTRG_3D(T), tk.truncrank(8), tk.maxiter(25), verbosity=0)
        assert rel(tk.free_energy(data, tk.ising_βc_3D, scalefactor=8.0), -3.508) < 1.0e-3


def test_two_by_two_finalizer(tk):
    # finalize_two_by_two! (src/utility/finalize.jl:17-25) against the oracle
    T = tk.classical_ising(tk.Trivial, 0.42)
    got = tk.run(tk.TRG(T), tk.truncrank(8), tk.maxiter(5), tk.two_by_two_Finalizer, verbosity=0)
    s = o.TRG(T)
    ref = [o.finalize_two_by_two(s)]
    for _ in range(5):
        s.step(8)
        ref.append(o.finalize_two_by_two(s))
    assert np.max(np.abs(np.array(got) - ref) / np.abs(ref)) <= 1e-10
