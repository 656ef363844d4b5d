"""More rows of the reference's model testset (test/models.jl:5-26, 40-46: `TRG(model)`,
truncrank(16), maxiter(25), rtol 1e-3) through the product path on the GPU -- clock (Trivial, ZN),
six-vertex (Trivial, U1), real phi^4 (Trivial, Z2) -- and the block-sparse steps on U(1) / Z3 /
Z2 sectors against the oracle.  Device twin of the `u1` / `models` tests in
tests/test_host_sequencing_emulated.py and of tests/test_oracle_golden.py::test_models_golden_*."""
import math

import numpy as np
import pytest

import tnr_oracle as o

pytestmark = pytest.mark.gpu
RTOL = 1e-10
_SQ3 = 2.0 * math.log(math.sqrt(3.0) + 1.0) / 3.0
_SQ2 = math.log(math.sqrt(2.0) + 1.0)

MODELS = {
    "clock3": (lambda tk: tk.classical_clock(tk.Trivial, 3, _SQ3), _SQ3, -4.17924244901635, 1e-3),
    "clock3_z3": (lambda tk: tk.classical_clock(tk.ZNIrrep[3], 3, _SQ3), _SQ3, -4.17924244901635, 1e-3),
    "clock4": (lambda tk: tk.classical_clock(tk.Trivial, 4, _SQ2), _SQ2, 2 * o.f_onsager, 1e-3),
    "clock4_z4": (lambda tk: tk.classical_clock(tk.ZNIrrep[4], 4, _SQ2), _SQ2, 2 * o.f_onsager, 1e-3),
    "sixvertex": (lambda tk: tk.sixvertex(tk.Trivial), 1.0, 1.5 * math.log(0.75), 1e-3),
    "sixvertex_u1": (lambda tk: tk.sixvertex(tk.U1Irrep), 1.0, 1.5 * math.log(0.75), 1e-3),
    # values recorded from the reference itself ("This is an approximation!"): the LAPACK oracle
    # reproduces them to 1e-13 (tests/test_oracle_golden.py); on the device the late RG steps may
    # order (near-)degenerate singular values differently (weight 2^-i in f), hence 1e-6 here
    "phi4_real": (lambda tk: tk.phi4_real(tk.Trivial, 10, -1.0, 1.0), -1.0, 0.4241912271276211, 1e-6),
    "phi4_real_z2": (lambda tk: tk.phi4_real(10, -1.0, 1.0), -1.0, 0.4232381701937374, 1e-6),
    "phi4_complex": (lambda tk: tk.phi4_complex(tk.Trivial, 6, -1.0, 1.0), -1.0, 0.7583605364656325, 1e-6),
    "phi4_complex_u1": (lambda tk: tk.phi4_complex(6, -1.0, 1.0), -1.0, 0.7673189874157453, 1e-6),
    # test/models.jl:19 (commented out there, "approximation"): 13 one-dimensional U(1) sectors
    "xy_u1": (lambda tk: tk.classical_XY(tk.U1Irrep, 0.89351, 6), 0.89351, -1.0251, 1e-3),
}


@pytest.mark.parametrize("model", list(MODELS))
def test_models_testset_more_rows(tk, model):
    make, beta, answer, tol = MODELS[model]
    T = make(tk)
    s = tk.TRG(T)
    assert s.sym == (getattr(T, "charges", None) is not None)
    data = tk.run(s, tk.truncrank(16), tk.maxiter(25), verbosity=0)
    assert abs((tk.free_energy(data, beta) - answer) / answer) < tol


@pytest.mark.parametrize("name,chi,n", [("TRG", 16, 6), ("BTRG", 16, 6), ("HOTRG", 8, 4), ("ATRG", 12, 3)])
@pytest.mark.parametrize("model", ["sixvertex_u1", "clock3_z3", "phi4_real_z2"])
def test_block_sparse_u1_and_more_models_match_oracle(tk, ctx, name, chi, n, model):
    T = MODELS[model][0](tk)
    s = getattr(tk, name)(T)
    assert s.sym and s.T.N == T.N
    ctx.reset_counters()
    got = np.array(tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0))
    ref = np.array(o.run(getattr(o, name)(np.asarray(T)), chi, n))
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL
    assert ctx.counters()["grouped_gemm_launches"] > 0
    if model == "sixvertex_u1":
        for key in s.T.blocks:      # U(1): exact conservation, no modulus
            assert sum(l.sign * q for l, q in zip(s.T.legs, key)) == 0


@pytest.mark.parametrize("chi", [16, 13])
@pytest.mark.parametrize("model", ["xy_u1", "sixvertex_u1", "phi4_complex_u1", "clock4_z4"])
def test_block_sparse_trg_equals_sector_oracle_on_device(tk, model, chi):
    """TensorKit's sector-global truncrank (oracle/sym_oracle.py) on tensors with exactly
    degenerate sectors, cuts through the multiplets included (chi = 13): gauge-invariant norm
    lists agree.  1e-10: the north star's tolerance (CPU twin: 1e-12)."""
    import sym_oracle as so

    T = tk.classical_clock(tk.ZNIrrep[4], 4, 0.88) if model == "clock4_z4" else MODELS[model][0](tk)
    ref = np.array(o.run(so.TRG_sym(np.asarray(T), T.charges, T.signs, T.N), chi, 10))
    got = np.array(tk.run(tk.TRG(T), tk.truncrank(chi), tk.maxiter(10), verbosity=0))
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL


@pytest.mark.parametrize("name", ["TRG", "BTRG", "HOTRG", "ATRG"])
@pytest.mark.parametrize("model", ["ising_trivial", "ising_z2", "sixvertex_u1"])
def test_spaces_testset(tk, name, model):
    """test/spaces.jl:5-24: every 2D scheme takes every kind of space through 25 steps at the odd
    truncrank(7) without a space mismatch (Gross-Neveu is fermionic: out of scope).  CPU twin:
    tests/test_host_sequencing_emulated.py::test_emulated_spaces_testset."""
    T = {"ising_trivial": lambda: tk.classical_ising(tk.Trivial),
         "ising_z2": lambda: tk.classical_ising(),
         "sixvertex_u1": lambda: tk.sixvertex(tk.U1Irrep)}[model]()
    s = getattr(tk, name)(T)
    assert s.sym == (model != "ising_trivial")
    data = tk.run(s, tk.truncrank(7), tk.maxiter(25), verbosity=0)
    assert len(data) == 26 and all(np.isfinite(x) and x > 0 for x in data)
    assert max(s.T.dims) <= 7
    if s.sym:
        for key in s.T.blocks:
            assert s.T.allowed(key)
