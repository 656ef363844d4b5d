"""CPU tests of the block-sparse HOST SEQUENCING (tnrkit.jl_b200/symmetric.py) with the C-ABI
primitives executed by tests/abi_emulator.py instead of libtnrcuda.  What is checked here is
the sector bookkeeping -- which blocks exist, arrows, offsets, leg orders, sector-global
truncation, chunking -- against the dense / sector-aware oracle; the arithmetic of the real
primitives is checked by the `-m gpu` suite, which runs the same sequences on the device."""
import itertools

import numpy as np
import pytest

import tnr_oracle as o
from abi_emulator import EmulatedContext

RTOL = 1e-10


@pytest.fixture()
def emu(tk, monkeypatch):
    """Routes every `ctx.call` of the host layer to the numpy emulation (tests only)."""
    from tnrkit.jl_b200 import _lib

    ctx = EmulatedContext()
    monkeypatch.setattr(_lib, "_default_ctx", ctx)
    return ctx


def _random_sym(rng, N, legs):
    dims = tuple(l.total for l in legs)
    a = rng.standard_normal(dims)
    q = [np.concatenate([[c] * l.dims[c] for c in l.charges]) for l in legs]
    for idx in itertools.product(*[range(d) for d in dims]):
        if sum(l.sign * q[i][j] for i, (l, j) in enumerate(zip(legs, idx))) % N != 0:
            a[idx] = 0.0
    return a


@pytest.mark.parametrize("N", [2, 3])
def test_emulated_contract_and_svd(tk, emu, N):
    rng = np.random.default_rng(N)
    L = lambda s: tk.Leg({c: 2 + c for c in range(N)}, s)
    la = [L(+1), L(+1), L(-1), L(-1)]
    a = _random_sym(rng, N, la)
    A = tk.SymTensor.from_dense(a, N, la)
    assert np.array_equal(A.to_dense(), a)
    assert np.array_equal(A.permute((2, 0, 3, 1)).to_dense(), np.transpose(a, (2, 0, 3, 1)))
    lb = [L(+1), L(+1), L(-1)]
    b = _random_sym(rng, N, lb)
    B = tk.SymTensor.from_dense(b, N, lb)
    Cc = tk.sym_contract(A, "abxy", B, "xyc", "cab")
    assert emu.calls["tnr_gemm_grouped"] == 1
    assert np.abs(Cc.to_dense() - np.einsum("abxy,xyc->cab", a, b)).max() <= 1e-12
    U, S, V, eps = tk.sym_svd_trunc(A, 2, 5)
    n = la[0].total
    sref = np.linalg.svd(a.reshape(n * n, n * n), compute_uv=False)
    got = np.sort(np.concatenate([s.to_numpy() for s in S.values()]))[::-1]
    assert np.abs(got - sref[:5]).max() <= 1e-12 * sref[0]
    assert abs(eps - np.linalg.norm(sref[5:])) <= 1e-11 * sref[0]


@pytest.mark.parametrize("name,chi,n", [("TRG", 8, 6), ("BTRG", 8, 6), ("HOTRG", 6, 4), ("ATRG", 8, 4)])
@pytest.mark.parametrize("model", ["ising_z2", "potts_z3"])
def test_emulated_2d_block_sparse_schemes_match_oracle(tk, emu, name, chi, n, model):
    """The sequences the GPU suite verifies on the device reproduce the oracle on the emulation
    too: pins the emulation itself (tests/test_gpu_symmetric.py is the device twin)."""
    T = tk.classical_ising() if model == "ising_z2" else tk.classical_potts(3)
    s = getattr(tk, name)(T)
    assert s.sym and s.ctx is emu
    got = np.array(tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0))
    ref = np.array(o.run(getattr(o, name)(np.asarray(T)), chi, n))
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL
    assert s.T.nnz() < np.prod(s.T.dims)


def _model(tk, which):
    import math

    if which == "sixvertex_u1":
        return tk.sixvertex(tk.U1Irrep), 1.0, 1.5 * math.log(0.75), 1e-3
    if which == "clock3_z3":
        b = 2.0 * math.log(math.sqrt(3.0) + 1.0) / 3.0
        return tk.classical_clock(tk.ZNIrrep[3], 3, b), b, -4.17924244901635, 1e-3
    if which == "phi4_z2":
        return tk.phi4_real(10, -1.0, 1.0), -1.0, 0.4232381701937374, 1e-11
    if which == "phi4_complex_u1":     # 36-dimensional legs, 11 U(1) sectors, 891 blocks
        return tk.phi4_complex(6, -1.0, 1.0), -1.0, 0.7673189874157453, 1e-10
    if which == "xy_u1":               # 13 one-dimensional U(1) sectors -6..6 (test/models.jl:19,
        # commented out there as "approximation": -1.0251); +-q sectors are exactly degenerate
        return tk.classical_XY(tk.U1Irrep, 0.89351, 6), 0.89351, -1.0251, 1e-3
    raise KeyError(which)


@pytest.mark.parametrize("name,chi,n", [("TRG", 16, 6), ("BTRG", 16, 6), ("HOTRG", 8, 4), ("ATRG", 12, 3)])
@pytest.mark.parametrize("model", ["sixvertex_u1", "clock3_z3", "phi4_z2"])
def test_emulated_block_sparse_u1_and_more_models(tk, emu, name, chi, n, model):
    """U(1) sectors (six-vertex: charges +-1/2 stored doubled, no modulus), Z3 clock and the
    Z2 phi^4 tensor with 5-dimensional sectors: block-sparse steps == dense oracle at 1e-10."""
    T = _model(tk, model)[0]
    s = getattr(tk, name)(T)
    assert s.sym and s.T.N == T.N
    got = np.array(tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0))
    ref = np.array(o.run(getattr(o, name)(np.asarray(T)), chi, n))
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL
    assert s.T.nnz() < np.prod(s.T.dims)
    if model == "sixvertex_u1":
        assert "U1" in repr(s.T)
        # U(1): the charges on the coarse legs grow with the bond dimension (no wrap-around)
        assert max(abs(q) for q in s.T.legs[0].charges) > 1
        for key in s.T.blocks:
            assert sum(l.sign * q for l, q in zip(s.T.legs, key)) == 0


@pytest.mark.parametrize("model", ["sixvertex_u1", "clock3_z3", "phi4_z2", "phi4_complex_u1", "xy_u1"])
def test_emulated_models_testset_block_sparse(tk, emu, model):
    """test/models.jl:40-46 on the symmetric tensors, block-sparse: TRG, truncrank(16), maxiter(25)."""
    T, beta, answer, tol = _model(tk, model)
    data = tk.run(tk.TRG(T), tk.truncrank(16), tk.maxiter(25), verbosity=0)
    f = tk.free_energy(data, beta)
    assert abs((f - answer) / answer) < tol


@pytest.mark.parametrize("chi", [16, 13])
@pytest.mark.parametrize("model", ["xy_u1", "sixvertex_u1", "phi4_complex_u1", "clock4_z4"])
def test_block_sparse_trg_equals_sector_oracle_with_degenerate_sectors(tk, emu, model, chi):
    """TensorKit semantics (oracle/sym_oracle.py: per-sector SVD, chi largest values over all
    sectors) on tensors whose +-q (U(1)) or clock sectors are EXACTLY degenerate, with cuts through
    the multiplets (chi = 13): the dense charge-basis oracle is not a valid reference there (its
    SVD mixes the degenerate sectors), the sector oracle is, and the block-sparse TRG agrees with
    it to rounding over 10 RG steps."""
    import sym_oracle as so

    T = tk.classical_clock(tk.ZNIrrep[4], 4, 0.88) if model == "clock4_z4" else _model(tk, model)[0]
    ref = np.array(o.run(so.TRG_sym(np.asarray(T), T.charges, T.signs, T.N), chi, 10))
    got = np.array(tk.run(tk.TRG(T), tk.truncrank(chi), tk.maxiter(10), verbosity=0))
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= 1e-12


def test_u1_tensor_bookkeeping(tk, emu):
    """SymTensor with N = 0: conservation without a modulus, coupled sectors keyed by the sum."""
    legs = [tk.Leg({-1: 2, 0: 1, 2: 3}, +1), tk.Leg({-2: 1, 1: 2}, +1), tk.Leg({-1: 2, 0: 2, 3: 1}, -1)]
    rng = np.random.default_rng(9)
    dims = tuple(l.total for l in legs)
    a = rng.standard_normal(dims)
    q = [np.concatenate([[c] * l.dims[c] for c in l.charges]) for l in legs]
    for idx in itertools.product(*[range(d) for d in dims]):
        if sum(l.sign * q[i][j] for i, (l, j) in enumerate(zip(legs, idx))) != 0:
            a[idx] = 0.0
    A = tk.SymTensor.from_dense(a, 0, legs)
    assert set(A.blocks) == {(-1, 1, 0), (2, -2, 0), (2, 1, 3)}
    assert all(sum(l.sign * c for l, c in zip(legs, k)) == 0 for k in A.blocks)
    assert np.array_equal(A.to_dense(), a)
    U, S, V, _ = tk.sym_svd_trunc(A, 2, 100)
    sref = np.linalg.svd(a.reshape(dims[0] * dims[1], dims[2]), compute_uv=False)
    got = np.sort(np.concatenate([s.to_numpy() for s in S.values()]))[::-1]
    assert np.abs(got - sref[:len(got)]).max() <= 1e-12 * sref[0]
    back = tk.sym_contract(U.scale_leg(2, S), "abk", V, "kc", "abc")
    assert np.abs(back.to_dense() - a).max() <= 1e-12
    with pytest.raises(ValueError):
        tk.SymTensor(1, legs, {})


@pytest.mark.parametrize("name,chi,n", [("HOTRG_3D", 4, 3), ("HOTRG_3D", 6, 2), ("ATRG_3D", 4, 3),
                                        ("ATRG_3D", 6, 2)])
def test_emulated_3d_block_sparse_schemes_match_oracle(tk, emu, name, chi, n):
    """HOTRG_3D / ATRG_3D on the Z2 tensor the reference's 3D testsets use (test/schemes.jl:8)."""
    T = tk.classical_ising_3D()
    s = getattr(tk, name)(T, symmetric=True)
    assert s.sym and s.ctx is emu
    got = np.array(tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0))
    ref = np.array(o.run(getattr(o, name)(np.asarray(T)), chi, n))
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL
    assert s.T.nnz() < np.prod(s.T.dims)
    # arrows: HOTRG_3D reproduces the input spaces; ATRG_3D's new bonds are non-dual in the domain
    # of H and in the codomain of Proj_2/4 (atrg3d.jl:68-82), which after its leg rotations gives
    # S' (x) S <- S' (x) S' (x) S (x) S from the second step on
    want = (1, -1, -1, -1, 1, 1) if name == "HOTRG_3D" else (-1, 1, 1, 1, -1, -1)
    assert tuple(l.sign for l in s.T.legs) == want


def test_emulated_hotrg3d_chunking_is_exact(tk, emu):
    """Chunking the two open x-bonds (ragged chunks, cached and recomputed P_D) changes nothing."""
    from tnrkit.jl_b200 import symmetric

    T = tk.classical_ising_3D()
    chi, n = 6, 3
    base = np.array(tk.run(tk.HOTRG_3D(T, symmetric=True), tk.truncrank(chi), tk.maxiter(n),
                           verbosity=0))
    assert symmetric.LAST_PLAN["hotrg3d"]["chunks"] == 1
    seen = set()
    for budget in (10, 6 ** 6 * 2, 6 ** 6 * 3):
        s = tk.HOTRG_3D(T, symmetric=True, max_chunk_elems=budget)
        got = np.array(tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0))
        plan = symmetric.LAST_PLAN["hotrg3d"]
        assert plan["chunks"] > 1
        seen.add((plan["chunk_size"], plan["cached_P"]))
        assert np.max(np.abs(got - base) / np.abs(base)) <= 1e-13
    assert any(c for _, c in seen) and any(not c for _, c in seen), seen
    assert any(sz == 2 for sz, _ in seen), seen     # 3-dimensional sectors in chunks of 2: ragged


def test_sym_slice_and_scatter_roundtrip(tk, emu):
    rng = np.random.default_rng(5)
    legs = [tk.Leg({0: 3, 1: 2}, +1), tk.Leg({0: 2, 1: 3}, -1), tk.Leg({0: 5, 1: 4}, +1)]
    a = _random_sym(rng, 2, legs)
    A = tk.SymTensor.from_dense(a, 2, legs)
    from tnrkit.jl_b200.symmetric import leg_chunks, sym_scatter, sym_slice, sym_zeros

    out = sym_zeros(2, legs, emu)
    chunks = leg_chunks(legs[2], 2)
    assert chunks == [(0, 0, 2), (0, 2, 4), (0, 4, 5), (1, 0, 2), (1, 2, 4)]
    for ch in chunks:
        piece = sym_slice(A, 2, ch)
        q, lo, hi = ch
        off = legs[2].offsets[q]
        dense = piece.to_dense()
        assert dense.shape == (5, 5, hi - lo)
        assert np.array_equal(dense, a[:, :, off + lo: off + hi])
        sym_scatter(out, piece, {2: ch})
    assert np.array_equal(out.to_dense(), a)


@pytest.mark.parametrize("chi,n,chunk", [(4, 3, 1), (6, 2, 2), (6, 2, 100)])
def test_emulated_atrg3d_chunked_tail_matches_oracle(tk, emu, chi, n, chunk):
    """Block-sparse ATRG_3D with AX / YD never formed whole: chunks of their open bond, R factors
    by TSQR over the chunk R factors, H / G assembled chunk by chunk (the single-process form of
    the sharded step, symmetric.py: _atrg3d_tail_sharded).  Ragged chunks at chi=6 / chunk=2
    (3-dimensional sectors); chunk=100: one chunk per sector."""
    from tnrkit.jl_b200 import symmetric

    T = tk.classical_ising_3D()
    s = tk.ATRG_3D(T, symmetric=True, sym_chunk=chunk)
    got = np.array(tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0))
    ref = np.array(o.run(o.ATRG_3D(np.asarray(T)), chi, n))
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL
    plan = symmetric.LAST_PLAN["atrg3d"]
    assert plan["world"] == 1 and plan["chunks_AX"] >= 2 and plan["chunks_YD"] >= 2
    assert tuple(l.sign for l in s.T.legs) == (-1, 1, 1, 1, -1, -1)
    base = np.array(tk.run(tk.ATRG_3D(T, symmetric=True), tk.truncrank(chi), tk.maxiter(n),
                           verbosity=0))
    assert np.max(np.abs(got - base) / np.abs(base)) <= 1e-11


@pytest.mark.parametrize("name", ["TRG", "BTRG", "HOTRG", "ATRG"])
@pytest.mark.parametrize("model", ["ising_z2", "sixvertex_u1"])
def test_emulated_spaces_testset(tk, emu, name, model):
    """test/spaces.jl:5-24: every 2D scheme takes every kind of space through 25 steps at the odd
    truncrank(7) without a space mismatch (here: the Z2 Ising and the U(1) six-vertex tensor on
    the block-sparse path; `classical_ising(Trivial)` runs on the dense device path, GPU twin in
    tests/test_gpu_models_u1.py; the Gross-Neveu tensor is fermionic: out of scope).  On top
    of the reference's `isa(..., Any)`: all 26 norms are finite and positive, every block obeys
    the conservation law, and contracted bonds keep opposite arrows."""
    T = tk.classical_ising() if model == "ising_z2" else tk.sixvertex(tk.U1Irrep)
    s = getattr(tk, name)(T)
    assert s.sym
    data = tk.run(s, tk.truncrank(7), tk.maxiter(25), verbosity=0)
    assert len(data) == 26 and all(np.isfinite(x) and x > 0 for x in data)
    assert max(s.T.dims) <= 7
    for key in s.T.blocks:
        assert s.T.allowed(key)
    # horizontal / vertical bonds of the final tensor can still be contracted with themselves
    assert s.T.legs[0].sign == -s.T.legs[3].sign and s.T.legs[0].same_space(s.T.legs[3])
    assert s.T.legs[1].sign == -s.T.legs[2].sign and s.T.legs[1].same_space(s.T.legs[2])


def _oracle_twin(tk, name, s):
    """Oracle scheme object holding the same state as the product scheme `s`."""
    dense = s.T.to_dense() if s.sym else s.T.to_numpy()
    os_ = getattr(o, name)(dense)
    if name == "BTRG":
        if s.sym:
            w = lambda S, leg: np.concatenate([S[q].to_numpy().reshape(-1) for q in s.T.legs[leg].charges])
            os_.S1, os_.S2 = np.diag(w(s.S1, 1)), np.diag(w(s.S2, 0))
        else:
            os_.S1, os_.S2 = np.diag(s.S1.to_numpy().reshape(-1)), np.diag(s.S2.to_numpy().reshape(-1))
    return os_


@pytest.mark.parametrize("name", ["TRG", "BTRG", "HOTRG", "ATRG"])
def test_emulated_cft_observables_match_oracle(tk, emu, name):
    """cft_data / central_charge / ground_state_degeneracy / gu_wen_ratio (tnrkit.jl_b200/cft.py;
    reference: src/utility/cft.jl) and their finalizers on a block-sparse scheme after a few RG
    steps: same numbers as the oracle's restatement on the same state (unit cells 1 and 2)."""
    s = getattr(tk, name)(tk.classical_ising(o.ising_bc + 0.01))
    assert s.sym
    tk.run(s, tk.truncrank(8), tk.maxiter(6), verbosity=0)
    tw = _oracle_twin(tk, name, s)
    from tnrkit.jl_b200 import cft as cft_mod

    assert cft_mod._sym_ok(s)               # sector-native path: no dense contraction is called
    dense_calls = emu.calls.get("tnr_contract", 0)
    grouped = emu.calls["tnr_gemm_grouped"]
    for u in (1, 2):
        got, ref = tk.cft_data(s, unitcell=u), o.cft_data(tw, unitcell=u)
        k = min(6, len(ref))
        assert len(got) == len(ref) and np.abs(got[:k] - ref[:k]).max() <= 1e-9
        assert abs(tk.ground_state_degeneracy(s, u) - o.ground_state_degeneracy(tw, u)) <= 1e-10
    assert np.abs(np.array(tk.gu_wen_ratio(s)) - np.array(o.gu_wen_ratio(tw))).max() <= 1e-10
    assert abs(tk.central_charge(s, 1.7) - o.central_charge(tw, 1.7)) <= 1e-10
    assert emu.calls.get("tnr_contract", 0) == dense_calls and emu.calls["tnr_gemm_grouped"] > grouped
    # finalizer forms through run!: one entry per step plus the initial one
    s2 = getattr(tk, name)(tk.classical_ising(o.ising_bc + 0.01))
    data = tk.run(s2, tk.truncrank(8), tk.maxiter(3), tk.guwenratio_Finalizer, verbosity=0)
    assert len(data) == 4 and all(len(d) == 2 for d in data)
    tw2 = getattr(o, name)(o.classical_ising(o.ising_bc + 0.01))
    ref = [o.finalize_gu_wen_ratio(tw2)]
    for _ in range(3):
        tw2.step(8)
        ref.append(o.finalize_gu_wen_ratio(tw2))
    assert np.abs(np.array(data) - np.array(ref)).max() <= 1e-9
    assert abs(tk.finalize_groundstatedegeneracy(s2) - o.finalize_groundstatedegeneracy(tw2)) <= 1e-9
    assert abs(tk.finalize_central_charge(s2) - o.finalize_central_charge(tw2)) <= 1e-9


def test_emulated_cft_observables_dense_tensor(tk, emu):
    """The dense (`Trivial`) form of the same functions: the state of an oracle BTRG / TRG run is
    loaded into dense product schemes."""
    for name in ("TRG", "BTRG"):
        tw = getattr(o, name)(o.classical_ising(o.ising_bc - 0.01))
        o.run(tw, 8, 5)
        s = getattr(tk, name)(tk.classical_ising(tk.Trivial))
        assert not s.sym
        s.T = tk.DeviceTensor.from_numpy(tw.T, 2, emu)
        if name == "BTRG":
            s.S1 = tk.DeviceTensor.from_numpy(np.diag(tw.S1).copy(), 1, emu)
            s.S2 = tk.DeviceTensor.from_numpy(np.diag(tw.S2).copy(), 1, emu)
        assert np.abs(tk.cft_data(s)[:5] - o.cft_data(tw)[:5]).max() <= 1e-10
        assert np.abs(tk.cft_data(s, unitcell=3)[:5] - o.cft_data(tw, unitcell=3)[:5]).max() <= 1e-9
        assert abs(tk.ground_state_degeneracy(s) - o.ground_state_degeneracy(tw)) <= 1e-10
        assert np.abs(np.array(tk.gu_wen_ratio(s)) - np.array(o.gu_wen_ratio(tw))).max() <= 1e-10
        assert abs(tk.central_charge(s, 0.9) - o.central_charge(tw, 0.9)) <= 1e-10
    with pytest.raises(TypeError):
        tk.cft_data(tk.HOTRG_3D(tk.classical_ising_3D(tk.Trivial), shard=False))


@pytest.mark.parametrize("name", ["TRG", "BTRG", "HOTRG", "ATRG"])
def test_emulated_two_by_two_finalizer(tk, emu, name):
    """finalize_two_by_two! (finalize.jl:17-25; BTRG with its bond weights: 27-42) through run!
    on a block-sparse scheme, and on a dense scheme loaded with the same state."""
    s = getattr(tk, name)(tk.classical_ising())
    got = tk.run(s, tk.truncrank(8), tk.maxiter(4), tk.two_by_two_Finalizer, verbosity=0)
    tw = getattr(o, name)(o.classical_ising())
    ref = [o.finalize_two_by_two(tw)]
    for _ in range(4):
        tw.step(8)
        ref.append(o.finalize_two_by_two(tw))
    assert np.max(np.abs(np.array(got) - np.array(ref)) / np.abs(ref)) <= RTOL
    if name in ("TRG", "BTRG"):
        d = getattr(tk, name)(tk.classical_ising(tk.Trivial))
        tw.T = 1.7 * tw.T
        d.T = tk.DeviceTensor.from_numpy(tw.T, 2, emu)
        if name == "BTRG":
            d.S1 = tk.DeviceTensor.from_numpy(np.diag(tw.S1).copy(), 1, emu)
            d.S2 = tk.DeviceTensor.from_numpy(np.diag(tw.S2).copy(), 1, emu)
        want = o.finalize_two_by_two(tw)
        assert abs(want - 1.7) <= 1e-9
        assert abs(tk.finalize_two_by_two(d) - want) <= RTOL * want
        assert np.abs(d.T.to_numpy() - tw.T).max() <= 1e-12 * np.abs(tw.T).max()
