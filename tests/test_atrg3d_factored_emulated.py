"""CPU tests of the factored ATRG_3D step (tnrkit.jl_b200/atrg3d_factored.py): the host layer only
sequences C-ABI primitives, which tests/abi_emulator.py executes with numpy here (tests only; the
`-m gpu` twin tests/test_gpu_atrg3d_factored.py runs the same sequences on the device).

What is checked: the two-factor algebra (leg names, implicit products, trace), the truncated SVD
of an implicit operator against a dense LAPACK SVD, the whole step against the oracle's
restatement of atrg3d.jl at 1e-10, exactness of chunking, and the world-size-2 sharding (gloo)."""
import os
import socket
import sys

import numpy as np
import pytest

import tnr_oracle as o
from abi_emulator import EmulatedContext

RTOL = 1e-10
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def emu(tk, monkeypatch):
    from tnrkit.jl_b200 import _lib

    ctx = EmulatedContext()
    monkeypatch.setattr(_lib, "_default_ctx", ctx)
    return ctx


def _two_factor(tk, rng, dims, bond, decay=0.6):
    """Random T[a b c d e f] = sum_i P[a c d i] Q[b e f i] with geometrically decaying weights."""
    from tnrkit.jl_b200.atrg3d_factored import TwoFactor

    a, b, c, d, e, f = dims
    P = rng.standard_normal((a, c, d, bond)) * decay ** np.arange(bond)
    Q = rng.standard_normal((b, e, f, bond))
    F = TwoFactor(tk.DeviceTensor.from_numpy(P), "acdi", tk.DeviceTensor.from_numpy(Q), "befi",
                  "abcdef")
    return F, np.einsum("acdi,befi->abcdef", P, Q)


def test_two_factor_algebra(tk, emu):
    from tnrkit.jl_b200.atrg3d_factored import TwoFactor

    rng = np.random.default_rng(1)
    F, dense = _two_factor(tk, rng, (2, 3, 4, 3, 2, 5), 4)
    assert F.dims == dense.shape and F.bond_dim == 4
    assert np.abs(F.to_dense().to_numpy() - dense).max() <= 1e-13
    perm = (3, 5, 1, 4, 0, 2)
    G = F.permute(perm)
    assert G.dims == tuple(dense.shape[p] for p in perm)
    assert np.abs(G.to_dense().to_numpy() - np.transpose(dense, perm)).max() <= 1e-13
    R = G.relabel()
    assert R.legs == "abcdef" and R.dims == G.dims
    assert np.abs(R.to_dense().to_numpy() - np.transpose(dense, perm)).max() <= 1e-13
    # implicit products in both directions, for the matricization the step uses
    dp = np.transpose(dense, perm)
    V = rng.standard_normal((dp.shape[2], dp.shape[3], dp.shape[0], 3))
    got = R.apply("bef", "cda", tk.DeviceTensor.from_numpy(V)).to_numpy()
    assert np.abs(got - np.einsum("abcdef,cdak->befk", dp, V)).max() <= 1e-12
    Z = rng.standard_normal((dp.shape[1], dp.shape[4], dp.shape[5], 2))
    got = R.apply("cda", "bef", tk.DeviceTensor.from_numpy(Z)).to_numpy()
    assert np.abs(got - np.einsum("abcdef,befk->cdak", dp, Z)).max() <= 1e-12
    # exact factorization of an explicit tensor and the trace T[1 1; 2 3 2 3]
    sq = rng.standard_normal((3, 3, 2, 4, 2, 4))
    E = TwoFactor.from_dense(tk.DeviceTensor.from_numpy(sq, 2))
    assert np.abs(E.to_dense().to_numpy() - sq).max() <= 1e-13
    assert abs(E.trace_3d() - np.einsum("aabcbc->", sq)) <= 1e-12
    p2 = np.transpose(sq, perm)
    assert p2.shape[0] == p2.shape[1]
    assert abs(E.permute(perm).trace_3d() - np.einsum("aabcbc->", p2)) <= 1e-12
    E.scale(0.5)
    assert np.abs(E.to_dense().to_numpy() - 0.5 * sq).max() <= 1e-13


@pytest.mark.parametrize("d,block", [(4, None), (5, None), (5, 12)])
def test_svd_topk_factored_matches_dense_svd(tk, emu, d, block):
    """d=4: b = chi + 64 exceeds the 64 x 64 matrix, one dense SVD; d=5: 69 or 12 of 125 columns,
    the subspace iteration proper."""
    from tnrkit.jl_b200.atrg3d_factored import svd_topk_factored

    rng = np.random.default_rng(2)
    F, dense = _two_factor(tk, rng, (d,) * 6, 30, decay=0.75)
    chi, N = 5, d ** 3
    st = {}
    U, S, V = svd_topk_factored(F, "bef", "cda", chi, stats=st, block=block)
    A = np.transpose(dense, (1, 4, 5, 2, 3, 0)).reshape((N, N), order="F")
    u, s, vh = np.linalg.svd(A)
    assert st["dense"] == (d == 4)
    if d > 4:
        assert st["iterations"] >= 1 and st["residual"] <= 2e-12
    assert np.abs(S.to_numpy() - s[:chi]).max() <= 1e-12 * s[0]
    Um = U.to_numpy().reshape(N, chi, order="F")
    Vm = V.to_numpy().reshape(N, chi, order="F")
    assert np.abs(Um.T @ Um - np.eye(chi)).max() <= 1e-11
    # gauge-invariant comparison: the rank-chi approximation itself
    approx = (Um * S.to_numpy()) @ Vm.T
    assert np.abs(approx - (u[:, :chi] * s[:chi]) @ vh[:chi]).max() <= 1e-10 * s[0]


def test_svd_topk_factored_uncertified_iteration_falls_back_or_raises(tk, emu):
    """maxit = 1 cannot certify: small matrices are decomposed densely instead, large ones raise
    (the result never rests on an uncertified subspace)."""
    from tnrkit.jl_b200.atrg3d_factored import svd_topk_factored

    rng = np.random.default_rng(4)
    F, dense = _two_factor(tk, rng, (5,) * 6, 40, decay=0.97)
    st = {}
    _, S, _ = svd_topk_factored(F, "bef", "cda", 5, stats=st, block=6, maxit=1)
    assert st["dense"] and "certify" in st["why"]
    s = np.linalg.svd(np.transpose(dense, (1, 4, 5, 2, 3, 0)).reshape(125, 125), compute_uv=False)
    assert np.abs(S.to_numpy() - s[:5]).max() <= 1e-12 * s[0]
    with pytest.raises(tk.TNRCudaError):
        svd_topk_factored(F, "bef", "cda", 5, block=6, maxit=1, dense_fallback_elems=100)


def test_svd_topk_factored_rank_deficient_operator(tk, emu):
    """rank(A) = 3 < chi = 5: the missing triplets are returned as zeros (they enter every later
    contraction with weight sigma or sqrt(sigma))."""
    from tnrkit.jl_b200.atrg3d_factored import svd_topk_factored

    rng = np.random.default_rng(3)
    F, dense = _two_factor(tk, rng, (4,) * 6, 3, decay=0.5)
    st = {}
    U, S, V = svd_topk_factored(F, "acd", "bef", 5, stats=st, block=8)
    s = np.linalg.svd(np.transpose(dense, (0, 2, 3, 1, 4, 5)).reshape(64, 64), compute_uv=False)
    got = S.to_numpy()
    assert st["rank"] == 3 and got.shape == (5,)
    assert np.abs(got[:3] - s[:3]).max() <= 1e-12 * s[0] and np.all(got[3:] == 0.0)
    assert np.all(U.to_numpy()[..., 3:] == 0.0) and np.all(V.to_numpy()[..., 3:] == 0.0)


@pytest.mark.parametrize("chi,n,block", [(4, 3, None), (6, 3, None), (6, 3, 14), (10, 2, 24)])
def test_factored_atrg3d_matches_oracle(tk, emu, chi, n, block):
    """run!(ATRG_3D(T; factored), truncrank(chi), maxiter(n)) == the oracle's norm list at 1e-10.
    chi = 4: every SVD takes the dense branch; chi = 6: the iteration with the default block;
    block = 14 / 24: block ~ 2.4 chi as at chi = 48, where 2 chi + 16 << chi^3."""
    from tnrkit.jl_b200 import atrg3d_factored as af

    T = tk.classical_ising_3D(tk.Trivial)
    s = tk.ATRG_3D(T, factored=True, block=block)
    assert s.factors is not None and s.ctx is emu
    got = np.array(tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0))
    ref = np.array(o.run(o.ATRG_3D(T), chi, n))
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL
    assert "tnr_atrg3d_step" not in emu.calls and emu.calls["tnr_orth_r"] >= 4 * 3 * n
    if block is not None:
        assert all(not st["dense"] and st["iterations"] >= 2 for st in af.LAST_STATS["svd"])
    # the reference's field: T materialised on request, legs [D U; N E S W]
    assert s.T.dims == s.factors.dims == (chi,) * 6
    assert "factored" in repr(s)


def test_factored_atrg3d_matches_committed_golden_chi12(tk, emu):
    """tests/golden/oracle_norms.json: ATRG_3D at the reference's testset size chi = 12, 6 RG steps
    (block 76 of 1728 columns, chunked TSQR) -- the vector the device twin compares with too."""
    import json

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_norms.json")))
    ref = np.array(g["ATRG_3D_ising_trivial_chi12_it6"])
    s = tk.ATRG_3D(tk.classical_ising_3D(tk.Trivial), factored=True, max_chunk_elems=12 ** 5 * 5)
    data = tk.run(s, tk.truncrank(12), tk.maxiter(6), verbosity=0)
    got = np.array(data)
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL
    # the reference's own ATRG_3D testset (test/schemes.jl:365-373: truncrank(12), free energy
    # against f_benchmark3D = -3.507 at rtol 5e-3) on the same run: after 6 of its 25 iterations
    # the series sum_i log(z_i) 8^(1-i) is converged to 2e-5, and it must land on the value of the
    # oracle's full 25-step run (recorded from oracle/tnr_oracle.py, chi = 12) to that accuracy
    f = tk.free_energy(data, tk.ising_βc_3D, scalefactor=8.0)
    assert abs(f - (-3.507)) <= 5e-3 * 3.507
    assert abs(f - (-3.517692114222326)) <= 1e-4 * 3.5177


@pytest.mark.parametrize("chi,n,rfactor", [(6, 4, "gram"), (10, 3, "gram"), (6, 4, "gram_eigh")])
def test_factored_atrg3d_gram_r_factors_match_oracle(tk, emu, chi, n, rfactor):
    """rfactor="gram": R factors from the Gram matrices of the two-factor tensors (O(chi^6), no
    chunk is ever formed for them; pivoted Cholesky `tnr_psd_factor`, or the eigendecomposition
    with "gram_eigh") -- same norm lists as the oracle's Householder QR at 1e-10."""
    from tnrkit.jl_b200 import atrg3d_factored as af

    T = tk.classical_ising_3D(tk.Trivial)
    s = tk.ATRG_3D(T, factored=True, rfactor=rfactor)
    got = np.array(tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0))
    ref = np.array(o.run(o.ATRG_3D(T), chi, n))
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL
    assert af.LAST_STATS["rfactor"] == rfactor and "tnr_orth_r" not in emu.calls
    assert ("tnr_psd_factor" in emu.calls) == (rfactor == "gram")
    with pytest.raises(ValueError):
        tk.ATRG_3D(T, factored=True, rfactor="qr").step(tk.truncrank(4))


def test_factored_atrg3d_gram_matches_tsqr_vector_chi16(tk, emu):
    """chi = 16 (4096 x 4096 matricizations, 256 x 256 Gram matrices): the Gram path against the
    committed vector of the TSQR path (tests/golden/factored_cpu_norms.json)."""
    import json

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "factored_cpu_norms.json")))
    ref = np.array(g["ATRG_3D_ising_trivial_chi16_it3"])
    s = tk.ATRG_3D(tk.classical_ising_3D(tk.Trivial), factored=True, rfactor="gram")
    got = np.array(tk.run(s, tk.truncrank(16), tk.maxiter(3), verbosity=0))
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= 1e-11
    # and against the ORACLE's own chi = 16 run (tests/golden/baseline_sizes.json, generated by
    # oracle/tnr_oracle.py in 380 s of host time; conditioning-checked)
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "baseline_sizes.json")))
    g = g["ATRG_3D_ising_trivial_chi16_it4"]
    assert g["valid"]
    oracle = np.array(g["norms"][:4])
    assert np.max(np.abs(got - oracle) / np.abs(oracle)) <= RTOL


def test_factored_atrg3d_chunking_is_exact(tk, emu):
    """TSQR over chunks of the open bond (ragged chunks included) changes nothing."""
    from tnrkit.jl_b200 import atrg3d_factored as af

    T = tk.classical_ising_3D(tk.Trivial)
    chi, n = 5, 2
    base = np.array(tk.run(tk.ATRG_3D(T, factored=True), tk.truncrank(chi), tk.maxiter(n),
                           verbosity=0))
    assert af.LAST_STATS["chunks"]["AX"] == [1]
    ref = np.array(o.run(o.ATRG_3D(T), chi, n))
    assert np.max(np.abs(base - ref) / np.abs(ref)) <= RTOL
    for budget, nchunks in ((5 ** 5 * 2, 3), (1, 5)):
        s = tk.ATRG_3D(T, factored=True, max_chunk_elems=budget)
        got = np.array(tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0))
        assert af.LAST_STATS["chunks"]["AX"] == [nchunks] == af.LAST_STATS["chunks"]["YD"]
        assert np.max(np.abs(got - base) / np.abs(base)) <= 1e-12


def test_factored_is_chosen_when_dense_does_not_fit(tk, emu):
    T = tk.classical_ising_3D(tk.Trivial)
    s = tk.ATRG_3D(T)
    assert s.factored is None and s.factors is None
    assert not tk.ATRG_3D.wants_factored(24) and not tk.ATRG_3D.wants_factored(36)
    assert tk.ATRG_3D.wants_factored(40) and tk.ATRG_3D.wants_factored(48)
    with pytest.raises(ValueError):
        tk.ATRG_3D(T, factored=False, shard=True)


def test_chunk_plan_covers_the_bond():
    from tnrkit.jl_b200.atrg3d_factored import chunk_plan

    for n, world, width in ((48, 8, 1), (5, 2, 2), (7, 4, 3), (3, 4, 1)):
        seen = []
        for r in range(world):
            for lo, hi in chunk_plan(n, r, world, width):
                assert 0 < hi - lo <= width
                seen.extend(range(lo, hi))
        assert seen == list(range(n))


# ---------------------------------------------------------------------------------------
# world size 2 (gloo): chunks of the open bond divided between two processes
# ---------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, chi, n, budget, rfactor, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tnrkit.jl_b200 as tk
    from abi_emulator import EmulatedContext as Emu
    from tnrkit.jl_b200 import _lib, atrg3d_factored as af

    _lib._default_ctx = Emu()
    s = tk.ATRG_3D(tk.classical_ising_3D(tk.Trivial), shard=True, max_chunk_elems=budget,
                   rfactor=rfactor)
    assert s.factored and s.shard
    got = tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0)
    import json

    q.put((rank, got, dict(af.LAST_STATS["chunks"]), json.loads(json.dumps(af.LAST_STATS, default=str))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("chi,budget,rfactor", [(4, 1 << 28, "tsqr"),    # even split
                                                (5, 5 ** 5, "tsqr"),      # ragged 3 + 2, width 1
                                                (6, 2000, "gram")])       # dealt projector pairs,
#                                 sharded products of the subspace iterations (block 70 of 216 columns)
def test_factored_atrg3d_sharded_world2(chi, budget, rfactor):
    import torch.multiprocessing as mp

    n = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, chi, n, budget, rfactor, q))
             for r in range(2)]
    for p in procs:
        p.start()
    res = {r: (got, ch, st) for r, got, ch, st in (q.get(timeout=300) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import tnrkit.jl_b200 as tk

    ref = np.array(o.run(o.ATRG_3D(tk.classical_ising_3D(tk.Trivial)), chi, n))
    for r in range(2):
        got = np.array(res[r][0])
        assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL, r
        assert res[r][1]["world"] == 2 and len(res[r][1]["AX"]) == 2
        if rfactor == "gram":
            assert res[r][2]["projector_owners"] == [0, 1]
            assert all(not st["dense"] and st["cheap_iterations"] >= 1 for st in res[r][2]["svd"])
    assert res[0][0] == res[1][0]          # replicas stay bit-identical
    if chi == 5:
        assert res[0][1]["AX"] == [3, 2]   # ragged ownership of the open bond (3 + 2), width 1
