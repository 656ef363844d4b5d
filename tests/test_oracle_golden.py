"""Pins the CPU oracle to every golden number the reference holds for the hot path.

README.md:76-81        BTRG, truncrank(16), maxiter(25) at ising_βc: f = -2.1096504926141826902647832
test/schemes.jl:26     TRG   chi=24 it=25  rtol 2e-6 vs f_onsager
test/schemes.jl:65     BTRG  chi=24 it=25  rtol 6e-8
test/schemes.jl:103    HOTRG chi=16 it=25  rtol 6e-7 (scalefactor 4)
test/schemes.jl:141    ATRG  chi=24 it=25  rtol 3e-6 (scalefactor 4)
test/schemes.jl:360    ATRG_3D  chi=12 it=25 rtol 5e-3 vs -3.507 (scalefactor 8)
test/schemes.jl:370    HOTRG_3D chi=8  it=25 rtol 1e-3 vs -3.507
(all of them on the Z2-symmetric model; the dense charge-basis tensor reproduces it.)
test/models.jl:5-26,40-46  TRG chi=16 it=25, rtol 1e-3, one golden free energy per model: clock
                       (q = 3, 4; Trivial, ZN), six-vertex (Trivial, U1), real phi^4 (Trivial, Z2),
                       complex phi^4 (Trivial, U1)
                       -- the phi^4 values were recorded from the reference itself and are
                       reproduced to 1e-12 by oracle TRG + host model constructors.
"""
import math

import numpy as np
import pytest

import tnr_oracle as o


def rel(a, b):
    return abs((a - b) / b)


def test_readme_quickstart_golden():
    d = o.run(o.BTRG(o.classical_ising_z2basis()), 16, 25)
    assert len(d) == 26
    f = o.free_energy(d, o.ising_bc)
    assert rel(f, -2.1096504926141826902647832) < 1e-13
    assert abs(rel(f, o.f_onsager) - 3.1e-7) < 0.05e-7
    # the Trivial (no symmetry) tensor is the same network in another gauge
    f2 = o.free_energy(o.run(o.BTRG(o.classical_ising()), 16, 25), o.ising_bc)
    assert rel(f2, f) < 1e-12


@pytest.mark.parametrize("cls,chi,sf,tol", [(o.TRG, 24, 2.0, 2e-6), (o.BTRG, 24, 2.0, 6e-8),
                                            (o.HOTRG, 16, 4.0, 6e-7), (o.ATRG, 24, 4.0, 3e-6)])
def test_2d_schemes_reference_tolerances(cls, chi, sf, tol):
    d = o.run(cls(o.classical_ising_z2basis()), chi, 25)
    assert rel(o.free_energy(d, o.ising_bc, scalefactor=sf), o.f_onsager) < tol


def test_hotrg3d_reference_tolerance():
    d = o.run(o.HOTRG_3D(o.classical_ising_3D_z2basis()), 8, 25)
    assert rel(o.free_energy(d, o.ising_bc_3D, scalefactor=8.0), o.f_benchmark3D) < 1e-3


def test_atrg3d_reference_tolerance():
    # the reference runs 25 iterations (218 s here); the series sum_i log(z_i) 8^(1-i) has
    # converged to 1e-5 relative after 5, which is far inside the 5e-3 tolerance.
    d = o.run(o.ATRG_3D(o.classical_ising_3D_z2basis()), 12, 5)
    assert rel(o.free_energy(d, o.ising_bc_3D, scalefactor=8.0), o.f_benchmark3D) < 5e-3


def test_models_potts_3state():
    # test/models.jl:16: TRG chi=16 it=25 on classical_potts(Trivial, 3): -4.119552029995684, rtol 1e-3
    d = o.run(o.TRG(o.classical_potts(3)), 16, 25)
    assert rel(o.free_energy(d, o.potts_bc(3)), -4.119552029995684) < 1e-3


_SQ3 = 2.0 * math.log(math.sqrt(3.0) + 1.0) / 3.0
_SQ2 = math.log(math.sqrt(2.0) + 1.0)
MODEL_GOLDEN = [   # (name, constructor(tk), beta, reference value, tolerance)   test/models.jl:5-26
    ("clock3", lambda tk: tk.classical_clock(tk.Trivial, 3, _SQ3), _SQ3, -4.17924244901635, 1e-3),
    ("clock3_Z3", lambda tk: tk.classical_clock(tk.ZNIrrep[3], 3, _SQ3), _SQ3, -4.17924244901635, 1e-3),
    ("clock4", lambda tk: tk.classical_clock(tk.Trivial, 4, _SQ2), _SQ2, 2 * o.f_onsager, 1e-3),
    ("clock4_Z4", lambda tk: tk.classical_clock(tk.ZNIrrep[4], 4, _SQ2), _SQ2, 2 * o.f_onsager, 1e-3),
    ("sixvertex", lambda tk: tk.sixvertex(tk.Trivial), 1.0, 1.5 * math.log(0.75), 1e-3),
    ("sixvertex_U1", lambda tk: tk.sixvertex(tk.U1Irrep), 1.0, 1.5 * math.log(0.75), 1e-3),
    # "This is an approximation!" values = the reference's own output: reproduced to 1e-12
    ("phi4_real", lambda tk: tk.phi4_real(tk.Trivial, 10, -1.0, 1.0), -1.0, 0.4241912271276211, 1e-12),
    ("phi4_real_Z2", lambda tk: tk.phi4_real(10, -1.0, 1.0), -1.0, 0.4232381701937374, 1e-12),
    # complex phi^4, bond dimension 36 (K = 6): also the reference's own output; the U(1) tensor has
    # 11 charge sectors -5..5.  (Trivial: 2e-10, one near-degenerate cut reacts to summation order)
    ("phi4_complex", lambda tk: tk.phi4_complex(tk.Trivial, 6, -1.0, 1.0), -1.0, 0.7583605364656325, 1e-8),
    ("phi4_complex_U1", lambda tk: tk.phi4_complex(6, -1.0, 1.0), -1.0, 0.7673189874157453, 1e-11),
]


@pytest.mark.parametrize("name,make,beta,answer,tol", MODEL_GOLDEN, ids=[m[0] for m in MODEL_GOLDEN])
def test_models_golden_free_energies(tk, name, make, beta, answer, tol):
    """test/models.jl:40-46: `TRG(model)`, truncrank(16), maxiter(25), free_energy(data, temp)."""
    T = make(tk)
    d = o.run(o.TRG(np.asarray(T)), 16, 25)
    assert rel(o.free_energy(d, beta), answer) < tol
    if getattr(T, "charges", None) is not None:      # symmetric variants really are symmetric
        q = [np.asarray(c) for c in T.charges]
        tot = sum(s * q[i].reshape([-1 if j == i else 1 for j in range(4)])
                  for i, s in enumerate(T.signs))
        forbidden = (tot % T.N != 0) if T.N else (tot != 0)
        assert np.abs(np.asarray(T)[forbidden]).max() <= 1e-13 * np.abs(np.asarray(T)).max()


def test_free_energy_and_driver_semantics():
    # run! pushes one norm before the first step (finalize_beginning) and one per step
    s = o.TRG(o.classical_ising())
    assert len(o.run(s, 4, 3)) == 4
    assert len(o.run(o.TRG(o.classical_ising()), 4, 3, finalize_beginning=False)) == 3
    # free_energy: lnz = sum_i log(z_i) * sf^(x - i), x = 1 - log(initial)/log(sf)
    assert np.isclose(o.free_energy([np.e, np.e], 2.0), -(1.0 + 0.5) / 2.0)
    assert np.isclose(o.free_energy([np.e], 1.0, scalefactor=4.0, initial_size=4.0), -0.25)


def test_sector_oracle_agrees_with_dense_oracle():
    """Per-sector SVD + sector-global truncrank (TensorKit semantics, oracle/sym_oracle.py) gives
    the same norm list as the dense SVD of the charge-basis tensor when no cut is degenerate."""
    import sym_oracle as so

    for T, N, q in ((o.classical_ising_z2basis(), 2, (0, 1)),):
        s = so.TRG_sym(T, [q] * 4, (1, 1, -1, -1), N)
        a = o.run(s, 8, 6)
        b = o.run(o.TRG(T), 8, 6)
        assert np.max(np.abs(np.array(a) - np.array(b)) / np.abs(b)) < 1e-11
        sp1, sp2 = s.last_spectra
        assert sum(len(v) for _, v in sp1) == 8 and sum(len(v) for _, v in sp2) == 8


def test_ozaki_model_digit_planes_are_error_free():
    """CPU model of the opt-in INT8 engine: the digit expansion is exact (8 planes reproduce a
    double to 2^-56 of the row maximum) and the emulated product is as accurate as DGEMM."""
    import ozaki_model as om

    rng = np.random.default_rng(3)
    X = rng.standard_normal((7, 300)) * np.exp(rng.uniform(-8, 8, size=(7, 300)))
    planes, scale = om.split(X, 8)
    assert all(np.abs(p).max() <= 64 for p in planes)
    rec = om.reconstruct(planes, scale)
    # residual after 8 planes: |rem| <= 0.5 in units of 2^(-6-49) of the row scale 2^e
    assert np.abs((rec - X) / scale[:, None]).max() <= 2.0 ** -56
    assert np.all(scale >= np.abs(X).max(axis=1)) and np.all(scale <= 2 * np.abs(X).max(axis=1))
    A = rng.standard_normal((40, 512)) * np.exp(rng.uniform(-9, 9, size=(40, 512)))
    B = rng.standard_normal((30, 512)) * np.exp(rng.uniform(-9, 9, size=(30, 1)))
    ref = A.astype(np.longdouble) @ B.T.astype(np.longdouble)
    scale_ab = np.abs(A).astype(np.longdouble) @ np.abs(B.T).astype(np.longdouble)
    e8 = float(np.max(np.abs(om.multiply(A, B, 8) - ref) / scale_ab))
    e7 = float(np.max(np.abs(om.multiply(A, B, 7) - ref) / scale_ab))
    ed = float(np.max(np.abs(A @ B.T - ref) / scale_ab))
    assert e8 <= 2e-15 and ed <= 2e-15 and e7 <= 3e-13 and e8 < e7


# ---- observables of the coarse-grained tensor (src/utility/cft.jl), pinned to test/schemes.jl ----
@pytest.mark.parametrize("name,chi,n,r1,r2", [("TRG", 24, 10, 2.0e-4, 1.0e-2), ("BTRG", 24, 10, 3.0e-4, 2.0e-2),
                                              ("HOTRG", 16, 4, 6.0e-4, 1.0e-2), ("ATRG", 24, 3, 1.0e-2, 1.0e-2)])
def test_oracle_cft_data_reference_testsets(name, chi, n, r1, r2):
    """`cft_data(scheme)[2:end]` against `ising_cft_exact` at the reference's own sizes and
    tolerances -- test/schemes.jl:25-31 (TRG), 63-69 (BTRG), 100-106 (HOTRG), 137-143 (ATRG)."""
    s = getattr(o, name)(o.classical_ising())
    o.run(s, chi, n)
    cft = o.cft_data(s)[1:]
    assert abs(cft[0] - o.ising_cft_exact[0]) <= r1 * o.ising_cft_exact[0]
    assert abs(cft[1] - o.ising_cft_exact[1]) <= r2 * o.ising_cft_exact[1]


@pytest.mark.parametrize("name,chi", [("TRG", 16), ("BTRG", 16), ("HOTRG", 12), ("ATRG", 16)])
@pytest.mark.parametrize("dbeta,want", [(-0.01, 1.0), (+0.01, 2.0)])
def test_oracle_gsd_and_gu_wen_reference_testsets(name, chi, dbeta, want):
    """`ground_state_degeneracy` and `gu_wen_ratio` on both sides of beta_c, rtol 1e-2 --
    test/schemes.jl:33-54, 71-90, 108-127, 145-164."""
    s = getattr(o, name)(o.classical_ising(o.ising_bc + dbeta))
    o.run(s, chi, 20)
    x1, x2 = o.gu_wen_ratio(s)
    for got in (o.ground_state_degeneracy(s), x1, x2):
        assert abs(got - want) <= 1.0e-2 * want
    # the finalizer forms (finalize.jl:143-179) leave T normalised and return the same numbers
    assert abs(o.finalize_groundstatedegeneracy(s) - want) <= 1.0e-2 * want
    assert abs(o.finalize_gu_wen_ratio(s)[0] - want) <= 1.0e-2 * want
