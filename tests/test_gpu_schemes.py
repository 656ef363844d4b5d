"""GPU parity of the scheme step!/finalize! bodies against the CPU oracle on the same
inputs.  Compared quantities are the gauge-invariant ones the north star names: the
per-iteration norm list returned by run!, and free_energy, to rtol 1e-10 (FP64)."""
import numpy as np
import pytest

import tnr_oracle as o

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _norms(tk, cls, ocls, T, chi, n, **kw):
    got = tk.run(cls(T, **kw), tk.truncrank(chi), tk.maxiter(n), verbosity=0)
    ref = o.run(ocls(T, **kw), chi, n)
    return np.array(got), np.array(ref)


@pytest.mark.parametrize("name,chi,n", [("TRG", 8, 6), ("BTRG", 8, 6), ("HOTRG", 6, 4),
                                        ("ATRG", 8, 4)])
@pytest.mark.parametrize("model", ["ising_z2", "ising_trivial", "potts3"])
def test_2d_norms_match_oracle(tk, name, chi, n, model):
    T = {"ising_z2": lambda: tk.classical_ising(),
         "ising_trivial": lambda: tk.classical_ising(tk.Trivial, 0.4, h=0.1),
         "potts3": lambda: tk.classical_potts(tk.Trivial, 3)}[model]()
    got, ref = _norms(tk, getattr(tk, name), getattr(o, name), T, chi, n)
    assert got.shape == ref.shape == (n + 1,)
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL
    sf = 2.0 if name in ("TRG", "BTRG") else 4.0
    f_got = tk.free_energy(got, 0.4, scalefactor=sf)
    f_ref = o.free_energy(ref, 0.4, scalefactor=sf)
    assert abs(f_got - f_ref) <= RTOL * abs(f_ref)


def test_btrg_readme_golden(tk):
    """README.md:76-81: BTRG, truncrank(16), maxiter(25) at ising_βc -> f = -2.10965049261418269."""
    data = tk.run(tk.BTRG(tk.classical_ising(tk.ising_βc)), tk.truncrank(16), tk.maxiter(25),
                  verbosity=0)
    assert len(data) == 26
    f = tk.free_energy(data, tk.ising_βc)
    assert abs(f - (-2.1096504926141826902647832)) <= 1e-10 * abs(f)
    assert abs((f - tk.f_onsager) / tk.f_onsager) < 3.2e-7


def test_hotrg3d_norms_match_oracle(tk):
    T = tk.classical_ising_3D(tk.Trivial)
    got, ref = _norms(tk, tk.HOTRG_3D, o.HOTRG_3D, T, 4, 3)
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL
    T = tk.classical_ising_3D()
    got, ref = _norms(tk, tk.HOTRG_3D, o.HOTRG_3D, T, 5, 2)
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL


def test_hotrg3d_substep_slices(tk, ctx):
    """Sharded entry point: computing slice blocks separately equals the one-shot result."""
    import ctypes as C

    from tnrkit.jl_b200 import _lib
    rng = np.random.default_rng(0)
    T = rng.standard_normal((3, 3, 4, 4, 4, 4))
    T = T + np.transpose(T, (1, 0, 4, 5, 2, 3))
    dT = tk.DeviceTensor.from_numpy(T)
    chi = 6
    od = (3, 3, 6, 6, 6, 6)
    full = tk.DeviceTensor.empty(od)
    parts = tk.DeviceTensor.empty(od)
    do = (C.c_int64 * 6)()
    ctx.call("tnr_hotrg3d_substep", dT.ptr, _lib.i64(T.shape), chi, full.ptr, do, 0, 6)
    assert tuple(do) == od
    for lo, hi in ((0, 2), (2, 5), (5, 6)):
        ctx.call("tnr_hotrg3d_substep", dT.ptr, _lib.i64(T.shape), chi, parts.ptr, do, lo, hi)
    a, b = full.to_numpy(), parts.to_numpy()
    assert np.abs(a - b).max() <= 1e-12 * np.abs(a).max()
    # and equals the oracle's z-compression up to the projector sign gauge: compare |T'| traces
    s = o.HOTRG_3D(T)
    s._step(chi)
    ref = s.T
    tr = lambda X: np.einsum("aabcbc->", X)
    assert abs(tr(a) - tr(ref)) <= 1e-10 * abs(tr(ref))
    assert abs(np.linalg.norm(a) - np.linalg.norm(ref)) <= 1e-10 * np.linalg.norm(ref)


def test_atrg3d_norms_match_oracle(tk):
    T = tk.classical_ising_3D()
    got, ref = _norms(tk, tk.ATRG_3D, o.ATRG_3D, T, 4, 3)
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL


def test_golden_fixture_norms(tk):
    """Committed oracle vectors (tests/golden/oracle_norms.json, made by make_golden.py)."""
    import json
    import os

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_norms.json")))
    cases = [("TRG_ising_z2_chi16_it12", tk.TRG, tk.classical_ising(), 16, 12),
             ("HOTRG_ising_z2_chi12_it8", tk.HOTRG, tk.classical_ising(), 12, 8),
             ("ATRG_ising_z2_chi12_it8", tk.ATRG, tk.classical_ising(), 12, 8),
             ("HOTRG_3D_ising_trivial_chi6_it4", tk.HOTRG_3D, tk.classical_ising_3D(tk.Trivial), 6, 4),
             ("ATRG_3D_ising_z2_chi6_it4", tk.ATRG_3D, tk.classical_ising_3D(), 6, 4)]
    for key, cls, T, chi, n in cases:
        got = np.array(tk.run(cls(T), tk.truncrank(chi), tk.maxiter(n), verbosity=0))
        ref = np.array(g[key])
        assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL, key


def test_hotrg_chi32_uses_subspace_eigh_and_matches_oracle(tk):
    """HOTRG at chi = 32: the projector Gram matrices are 1024 x 1024 and only 32 eigenpairs are
    kept, so eigh_trunc runs the block subspace iteration; norms must still match LAPACK."""
    import ctypes as C

    ctx = tk.default_context()
    v0 = C.c_double()
    ctx.call("tnr_get_counter", b"subspace_eigh", C.byref(v0))
    T = tk.classical_ising(tk.Trivial, 0.43)
    got, ref = _norms(tk, tk.HOTRG, o.HOTRG, T, 32, 4)
    v1 = C.c_double()
    ctx.call("tnr_get_counter", b"subspace_eigh", C.byref(v1))
    assert v1.value > v0.value, "subspace path was never taken"
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL


def test_trg_chi32_uses_subspace_svd_and_matches_oracle(tk):
    """TRG at chi = 32 on the dense path: 1024 x 1024 matrices, 32 triplets kept -> subspace SVD."""
    import ctypes as C

    ctx = tk.default_context()
    v0, v1 = C.c_double(), C.c_double()
    ctx.call("tnr_get_counter", b"subspace_svd", C.byref(v0))
    T = tk.classical_ising(tk.Trivial, 0.43)
    got, ref = _norms(tk, tk.TRG, o.TRG, T, 32, 8)
    ctx.call("tnr_get_counter", b"subspace_svd", C.byref(v1))
    assert v1.value > v0.value, "subspace SVD path was never taken"
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL


def test_hotrg3d_split_entries_and_windowed_d_loop_are_bit_identical(tk, ctx):
    """`tnr_hotrg3d_proj_half` x 4 + `tnr_hotrg3d_contract` (what a sharded run calls) and the
    d loop blocked into windows of absorbed operands (`hotrg3d_pk_budget_mb`, the O(chi^7)
    memory cap) reproduce `tnr_hotrg3d_substep` bit for bit."""
    T = tk.classical_ising_3D(tk.Trivial)
    base = tk.run(tk.HOTRG_3D(T, shard=False), tk.truncrank(6), tk.maxiter(4), verbosity=0)
    forced = tk.run(tk.HOTRG_3D(T, shard=False, split_projectors="force"), tk.truncrank(6),
                    tk.maxiter(4), verbosity=0)
    assert forced == base
    ctx.set_option("hotrg3d_pk_budget_mb", 1)      # 6^6 doubles = 0.37 MB per operand: 2 per window
    try:
        windowed = tk.run(tk.HOTRG_3D(T, shard=False), tk.truncrank(6), tk.maxiter(4), verbosity=0)
    finally:
        ctx.set_option("hotrg3d_pk_budget_mb", 49152)
    assert windowed == base
    ref = np.array(o.run(o.HOTRG_3D(np.asarray(T)), 6, 4))
    assert np.max(np.abs(np.array(base) - ref) / np.abs(ref)) <= 1e-10
