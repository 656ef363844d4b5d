"""GPU parity of the primitives (through the C ABI) against numpy on the same inputs."""
import ctypes as C
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _up(tk, a):
    return tk.DeviceTensor.from_numpy(a)


def _gemm(tk, ctx, ta, tb, m, n, k, rng, alpha=1.0, beta=0.0):
    A = rng.standard_normal((k, m) if ta == "T" else (m, k))
    B = rng.standard_normal((n, k) if tb == "T" else (k, n))
    C0 = rng.standard_normal((m, n))
    dA, dB, dC = _up(tk, A), _up(tk, B), _up(tk, C0)
    ctx.call("tnr_gemm", ta.encode(), tb.encode(), m, n, k, alpha, dA.ptr, A.shape[0], dB.ptr,
             B.shape[0], beta, dC.ptr, m)
    ref = alpha * (A.T if ta == "T" else A) @ (B.T if tb == "T" else B) + beta * C0
    got = dC.to_numpy()
    return got, ref


@pytest.mark.parametrize("ta,tb", list(itertools.product("NT", "NT")))
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (7, 5, 3), (128, 128, 16), (129, 131, 37),
                                   (300, 24, 577), (64, 200, 4096), (513, 259, 65)])
def test_gemm_layouts(tk, ctx, ta, tb, m, n, k):
    rng = np.random.default_rng(m * 1000 + n * 10 + k)
    got, ref = _gemm(tk, ctx, ta, tb, m, n, k, rng)
    assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()) * np.sqrt(k)


def test_gemm_alpha_beta_splitk(tk, ctx):
    rng = np.random.default_rng(5)
    got, ref = _gemm(tk, ctx, "T", "N", 96, 80, 20000, rng, alpha=-0.5, beta=2.0)
    assert np.abs(got - ref).max() <= 1e-10


def test_gemm_strided_batched(tk, ctx):
    rng = np.random.default_rng(6)
    nb, m, n, k = 37, 45, 24, 33
    A = rng.standard_normal((nb, k, m))  # each A_b stored (m x k) column major == (k, m) C order
    B = rng.standard_normal((n, k))      # (k x n) column major
    dA = tk.DeviceTensor.from_numpy(np.transpose(A, (2, 1, 0)))
    dB = tk.DeviceTensor.from_numpy(B.T)
    dC = tk.DeviceTensor.empty((m, n, nb))
    ctx.call("tnr_gemm_strided_batched", b"N", b"N", m, n, k, 1.0, dA.ptr, m, m * k, dB.ptr, k, 0,
             0.0, dC.ptr, m, m * n, nb)
    got = dC.to_numpy()  # [m, n, nb]
    for b in range(nb):
        ref = A[b].T @ B.T
        assert np.abs(got[:, :, b] - ref).max() <= 1e-12


@pytest.mark.parametrize("dims,perm", [
    ((5,), (0,)), ((4, 7), (1, 0)), ((33, 65), (1, 0)), ((3, 4, 5), (2, 0, 1)),
    ((2, 3, 4, 5), (1, 3, 0, 2)), ((6, 5, 4, 3, 2, 7), (5, 3, 1, 2, 0, 4)),
    ((4, 4, 8, 8, 8, 8), (0, 1, 3, 2, 5, 4)), ((8, 8, 8, 8, 8, 8), (1, 3, 5, 0, 2, 4)),
    ((24, 24, 24), (2, 1, 0)), ((1, 9, 1, 4), (3, 2, 1, 0)), ((40, 3, 40), (0, 2, 1)),
])
def test_permute(tk, dims, perm):
    rng = np.random.default_rng(1)
    a = rng.standard_normal(dims)
    got = _up(tk, a).permute(perm).to_numpy()
    assert np.array_equal(got, np.transpose(a, perm))  # bit exact: pure data movement


def test_contract(tk):
    rng = np.random.default_rng(2)
    a = rng.standard_normal((3, 4, 5, 6))
    b = rng.standard_normal((5, 7, 3))
    for lc in ("bdf", "fbd", "dfb"):
        got = tk.contract(_up(tk, a), "abcd", _up(tk, b), "cfa", lc).to_numpy()
        ref = np.einsum("abcd,cfa->" + lc, a, b)
        assert np.abs(got - ref).max() <= 1e-12
    # outer product and full-K-leading / trailing natural layouts
    x, y = rng.standard_normal((4, 3)), rng.standard_normal((5,))
    got = tk.contract(_up(tk, x), "ab", _up(tk, y), "c", "abc").to_numpy()
    assert np.abs(got - np.einsum("ab,c->abc", x, y)).max() <= 1e-14


@pytest.mark.parametrize("m,n,chi", [(4, 4, 4), (16, 16, 5), (36, 20, 12), (20, 36, 40),
                                     (144, 144, 16), (7, 1, 3)])
def test_svd_trunc(tk, m, n, chi):
    rng = np.random.default_rng(m + n)
    a = rng.standard_normal((m, n)) * np.logspace(0, -6, n)[None, :]
    U, S, Vt, eps = tk.svd_trunc(_up(tk, a), 1, chi)
    u, s, vt = U.to_numpy(), S.to_numpy(), Vt.to_numpy()
    sref = np.linalg.svd(a, compute_uv=False)
    k = min(chi, m, n)
    assert s.shape == (k,)
    assert np.abs(s - sref[:k]).max() <= 1e-12 * sref[0]                    # spectrum
    assert abs(eps - np.linalg.norm(sref[k:])) <= 1e-12 * sref[0]          # truncation error
    assert np.abs(u.T @ u - np.eye(k)).max() <= 1e-12                      # isometries
    assert np.abs(vt @ vt.T - np.eye(k)).max() <= 1e-12
    uu, ss, vv = np.linalg.svd(a, full_matrices=False)
    best = (uu[:, :k] * ss[:k]) @ vv[:k]
    assert np.abs((u * s) @ vt - best).max() <= 1e-11 * sref[0]            # gauge invariant


def test_svd_rank_deficient(tk):
    # the initial Ising tensor: 4x4 matrix of rank 2 (exact zero singular values)
    a = tk.classical_ising().reshape(4, 4)
    U, S, Vt, eps = tk.svd_trunc(_up(tk, a), 1, 16)
    u, s, vt = U.to_numpy(), S.to_numpy(), Vt.to_numpy()
    assert np.all(np.isfinite(u)) and np.all(np.isfinite(vt))
    assert np.abs((u * s) @ vt - a).max() <= 1e-14
    assert np.abs(s - np.linalg.svd(a, compute_uv=False)).max() <= 1e-14


@pytest.mark.parametrize("n,chi", [(4, 2), (16, 16), (64, 8), (256, 16)])
def test_eigh_trunc(tk, n, chi):
    rng = np.random.default_rng(n)
    g = rng.standard_normal((n, n)) * np.logspace(0, -5, n)[None, :]
    mm = g @ g.T - 0.01 * np.outer(g[:, 0], g[:, 0])  # mostly PSD, one direction shifted
    W, V, eps = tk.eigh_trunc(_up(tk, mm), chi)
    w, v = W.to_numpy(), V.to_numpy()
    wr = np.linalg.eigvalsh(mm)
    order = np.argsort(-np.abs(wr))
    k = min(chi, n)
    scale = np.abs(wr).max()
    assert np.abs(w - wr[order[:k]]).max() <= 1e-12 * scale
    assert abs(eps - np.linalg.norm(wr[order[k:]])) <= 1e-12 * scale
    assert np.abs(v.T @ v - np.eye(k)).max() <= 1e-12
    assert np.abs(mm @ v - v * w).max() <= 1e-11 * scale


@pytest.mark.parametrize("m,n,k,alpha,beta", [(1664, 1664, 96, 1.0, 0.0), (1601, 1555, 78, -0.5, 2.0),
                                              (2048, 1536, 1000, 1.0, 0.0), (1538, 1666, 64, 2.0, 1.0)])
def test_gemm_tma_path(tk, ctx, m, n, k, alpha, beta):
    """TN problems with >= 148 tiles take the TMA/mbarrier kernel (ragged edges are zero
    filled by the tensor map); the result must equal numpy and the cp.async kernel."""
    rng = np.random.default_rng(m + n + k)
    before = ctx.counters()["tma_gemm_launches"]
    got, ref = _gemm(tk, ctx, "T", "N", m, n, k, rng, alpha, beta)
    assert ctx.counters()["tma_gemm_launches"] == before + 1, "TMA kernel was not used"
    assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()) * np.sqrt(k)
    ctx.set_option("disable_tma", 1)
    try:
        rng = np.random.default_rng(m + n + k)
        got2, _ = _gemm(tk, ctx, "T", "N", m, n, k, rng, alpha, beta)
    finally:
        ctx.set_option("disable_tma", 0)
    assert ctx.counters()["tma_gemm_launches"] == before + 1
    assert np.abs(got - got2).max() <= 1e-12 * max(1.0, np.abs(ref).max())


def test_gemm_tma_fallback_on_odd_leading_dimension(tk, ctx):
    """K = 77 gives an odd leading dimension (row stride not a multiple of 16 bytes): TMA cannot
    describe it, the cp.async kernel takes over and the result is still right."""
    rng = np.random.default_rng(9)
    before = ctx.counters()["tma_gemm_launches"]
    got, ref = _gemm(tk, ctx, "T", "N", 1601, 1555, 77, rng)
    assert ctx.counters()["tma_gemm_launches"] == before
    assert np.abs(got - ref).max() <= 1e-11


def _counter(ctx, name):
    import ctypes as C
    v = C.c_double()
    ctx.call("tnr_get_counter", name.encode(), C.byref(v))
    return v.value


@pytest.mark.parametrize("n,chi,decay", [(1536, 32, 8), (2048, 64, 12), (1024, 16, 4)])
def test_eigh_trunc_subspace_path(tk, ctx, n, chi, decay):
    """n >= 1024 with chi << n: top-chi eigenpairs come from the GEMM-rich block subspace
    iteration; values, truncation error and the invariant subspace must match LAPACK."""
    rng = np.random.default_rng(n + chi)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.logspace(0, -decay, n)
    mm = (q * lam) @ q.T
    mm = 0.5 * (mm + mm.T)
    before = _counter(ctx, "subspace_eigh")
    W, V, eps = tk.eigh_trunc(tk.DeviceTensor.from_numpy(mm), chi)
    assert _counter(ctx, "subspace_eigh") == before + 1, "subspace solver not used / fell back"
    w, v = W.to_numpy(), V.to_numpy()
    wr, vr = np.linalg.eigh(mm)
    order = np.argsort(-np.abs(wr))
    assert np.abs(w - wr[order[:chi]]).max() <= 1e-12 * np.abs(wr).max()
    assert abs(eps - np.linalg.norm(wr[order[chi:]])) <= 1e-7 * np.abs(wr).max()
    assert np.abs(v.T @ v - np.eye(chi)).max() <= 1e-12
    pr = vr[:, order[:chi]]
    # projector onto the kept subspace (gauge invariant); conditioning ~ eps / gap
    gap = abs(wr[order[chi - 1]]) - abs(wr[order[chi]])
    assert np.abs(v @ v.T - pr @ pr.T).max() <= 1e-13 * np.abs(wr).max() / gap * 50
    # and the full-Jacobi path gives the same answer
    ctx.set_option("disable_subspace", 1)
    try:
        W2, V2, eps2 = tk.eigh_trunc(tk.DeviceTensor.from_numpy(mm), chi)
    finally:
        ctx.set_option("disable_subspace", 0)
    # the full decomposition multiplies O(n^2) rotations per sweep: its own rounding error is
    # ~ n * eps * sweeps (1e-12 at n = 1536), the subspace result is the more accurate of the two
    assert np.abs(W2.to_numpy() - w).max() <= 1e-11 * np.abs(wr).max()
    assert abs(eps2 - eps) <= 1e-7 * np.abs(wr).max()


@pytest.mark.parametrize("m,n,chi,decay", [(1536, 1536, 32, 8), (2048, 1100, 48, 10), (1100, 2300, 16, 5)])
def test_svd_trunc_subspace_path(tk, ctx, m, n, chi, decay):
    """min(m, n) >= 1024 with chi << min(m, n): the top-chi triplets come from the block subspace
    iteration (GEMM + thin Jacobi); spectrum, truncation error and the rank-chi approximation
    must match LAPACK, and the full-Jacobi path must agree."""
    rng = np.random.default_rng(m + n + chi)
    r = min(m, n)
    u0, _ = np.linalg.qr(rng.standard_normal((m, r)))
    v0, _ = np.linalg.qr(rng.standard_normal((n, r)))
    s0 = np.logspace(0, -decay, r)
    a = (u0 * s0) @ v0.T
    before = _counter(ctx, "subspace_svd")
    U, S, Vt, eps = tk.svd_trunc(tk.DeviceTensor.from_numpy(a), 1, chi)
    assert _counter(ctx, "subspace_svd") == before + 1, "subspace solver not used / fell back"
    u, s, vt = U.to_numpy(), S.to_numpy(), Vt.to_numpy()
    sref = np.linalg.svd(a, compute_uv=False)
    assert np.abs(s - sref[:chi]).max() <= 1e-12 * sref[0]
    assert abs(eps - np.linalg.norm(sref[chi:])) <= 1e-7 * sref[0]
    assert np.abs(u.T @ u - np.eye(chi)).max() <= 1e-12
    assert np.abs(vt @ vt.T - np.eye(chi)).max() <= 1e-12
    uu, ss, vv = np.linalg.svd(a, full_matrices=False)
    best = (uu[:, :chi] * ss[:chi]) @ vv[:chi]
    gap = sref[chi - 1] - sref[chi]
    assert np.abs((u * s) @ vt - best).max() <= 1e-13 * sref[0] ** 2 / gap * 50


@pytest.mark.parametrize("opt", ["disable_block_jacobi", "disable_precondition"])
def test_jacobi_variants_agree(tk, ctx, opt):
    """Shared-memory block rounds / Gram preconditioning of tall problems are pure speed-ups:
    switching them off must give the same factorisations."""
    rng = np.random.default_rng(11)
    tall = rng.standard_normal((2000, 96)) * np.logspace(0, -9, 96)[None, :]
    sq = rng.standard_normal((200, 200))
    sq = sq @ sq.T * 1e-3
    res = {}
    for flag in (0, 1):
        ctx.set_option(opt, flag)
        try:
            U, S, Vt, eps = tk.svd_trunc(_up(tk, tall), 1, 40)
            W, V, e2 = tk.eigh_trunc(_up(tk, sq), 30)
            res[flag] = (S.to_numpy(), (U.to_numpy() * S.to_numpy()) @ Vt.to_numpy(), eps,
                         W.to_numpy(), e2)
        finally:
            ctx.set_option(opt, 0)
    sref = np.linalg.svd(tall, compute_uv=False)
    for flag in (0, 1):
        assert np.abs(res[flag][0] - sref[:40]).max() <= 1e-12 * sref[0]
    assert np.abs(res[0][1] - res[1][1]).max() <= 1e-11 * sref[0]
    assert abs(res[0][2] - res[1][2]) <= 1e-12 * sref[0]
    assert np.abs(res[0][3] - res[1][3]).max() <= 1e-12 * np.abs(res[0][3]).max()


def _counter(ctx, name):
    v = C.c_double()
    ctx.call("tnr_get_counter", name.encode(), C.byref(v))
    return v.value


@pytest.mark.parametrize("ta,tb", [("T", "N"), ("N", "N"), ("T", "T")])
def test_gemm_grouped_all_layouts_and_tma_path(tk, ctx, ta, tb):
    """`tnr_gemm_grouped`: per-sector products in one launch.  With both operands K-contiguous
    ("T", "N") and >= 148 tiles the launch runs on the TMA + mbarrier kernel with one pair of
    tensor maps per sector in a device-side table (counter `tma_grouped_launches`); ragged m, n, k
    are zero-filled by the maps.  Other layouts and small launches use the cp.async kernel."""
    from tnrkit.jl_b200 import _lib

    rng = np.random.default_rng(42)
    shapes = [(1500, 1300, 700), (1024, 1024, 1024), (130, 2000, 66), (777, 129, 4100), (8, 8, 64)]
    keep, probs, refs = [], [], []
    for m, n, k in shapes:
        A = rng.standard_normal((k, m) if ta == "T" else (m, k))
        B = rng.standard_normal((n, k) if tb == "T" else (k, n))
        C0 = rng.standard_normal((m, n))
        dA, dB, dC = _up(tk, A), _up(tk, B), _up(tk, C0)
        keep.append((dA, dB, dC))
        probs.append(_lib.GemmProblem(m, n, k, dA.buf.data_ptr(), A.shape[0], dB.buf.data_ptr(),
                                      B.shape[0], dC.buf.data_ptr(), m))
        refs.append(-0.5 * (A.T if ta == "T" else A) @ (B.T if tb == "T" else B) + 2.0 * C0)
    arr = (_lib.GemmProblem * len(probs))(*probs)
    t0, g0 = _counter(ctx, "tma_grouped_launches"), _counter(ctx, "grouped_gemm_launches")
    ctx.call("tnr_gemm_grouped", ta.encode(), tb.encode(), len(probs), arr, -0.5, 2.0)
    assert _counter(ctx, "grouped_gemm_launches") == g0 + 1
    assert _counter(ctx, "tma_grouped_launches") == t0 + (1 if (ta, tb) == ("T", "N") else 0)
    for (dA, dB, dC), ref, (m, n, k) in zip(keep, refs, shapes):
        got = dC.to_numpy()
        assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()) * np.sqrt(k), (m, n, k)


def test_gemm_grouped_tma_declines_small_or_unaligned(tk, ctx):
    from tnrkit.jl_b200 import _lib

    rng = np.random.default_rng(43)
    for shapes, odd_ld in [([(100, 90, 80), (64, 64, 64)], False), ([(1500, 1300, 701)] * 2, True)]:
        keep, probs, refs = [], [], []
        for m, n, k in shapes:
            A, B = rng.standard_normal((k, m)), rng.standard_normal((k, n))
            dA, dB, dC = _up(tk, A), _up(tk, B), tk.DeviceTensor.empty((m, n))
            keep.append((dA, dB, dC))
            probs.append(_lib.GemmProblem(m, n, k, dA.buf.data_ptr(), k, dB.buf.data_ptr(), k,
                                          dC.buf.data_ptr(), m))
            refs.append(A.T @ B)
        arr = (_lib.GemmProblem * len(probs))(*probs)
        t0 = _counter(ctx, "tma_grouped_launches")
        ctx.call("tnr_gemm_grouped", b"T", b"N", len(probs), arr, 1.0, 0.0)
        assert _counter(ctx, "tma_grouped_launches") == t0      # too few tiles / odd leading dim
        for (dA, dB, dC), ref in zip(keep, refs):
            assert np.abs(dC.to_numpy() - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("m,n", [(64, 32), (1000, 96), (5000, 100), (4097, 33), (300, 300), (20000, 256),
                                 (500, 7), (33, 1), (70000, 64), (2048, 577)])
def test_blocked_householder_qr(tk, ctx, m, n):
    """`tnr_qr` (qr.cu: panels of 32 Householder columns, compact WY, DMMA trailing update) against
    numpy / LAPACK geqrf: same R up to rounding INCLUDING signs (dlarfg convention), Q orthonormal
    to rounding, Q R = A."""
    rng = np.random.default_rng(m + n)
    A = rng.standard_normal((m, n))
    dA = _up(tk, A)
    dQ, dR = tk.DeviceTensor.empty((m, n)), tk.DeviceTensor.empty((n, n))
    q0 = _counter(ctx, "qr_factorizations")
    ctx.call("tnr_qr", dA.ptr, m, n, dQ.ptr, dR.ptr)
    assert _counter(ctx, "qr_factorizations") == q0 + 1
    Q, R = dQ.to_numpy(), dR.to_numpy()
    assert np.array_equal(np.tril(R, -1), np.zeros_like(R))
    scale = np.linalg.norm(A)
    assert np.abs(Q.T @ Q - np.eye(n)).max() <= 5e-14
    assert np.linalg.norm(Q @ R - A) <= 1e-14 * scale * np.sqrt(n)
    Rn = np.linalg.qr(A, mode="r")
    assert np.abs(R - Rn).max() <= 1e-12 * scale


def test_qr_rank_deficient_and_graded(tk, ctx):
    rng = np.random.default_rng(9)
    m, n = 3000, 80
    A = rng.standard_normal((m, n))
    A[:, 10] = 0.0                      # a zero column
    A[:, 20] = A[:, 5]                  # a duplicate
    A[:, 40:] *= 10.0 ** -np.arange(40)[None, :]     # graded over 40 orders of magnitude
    dA = _up(tk, A)
    dQ, dR = tk.DeviceTensor.empty((m, n)), tk.DeviceTensor.empty((n, n))
    ctx.call("tnr_qr", dA.ptr, m, n, dQ.ptr, dR.ptr)
    Q, R = dQ.to_numpy(), dR.to_numpy()
    assert np.isfinite(Q).all() and np.isfinite(R).all()
    assert np.abs(Q.T @ Q - np.eye(n)).max() <= 5e-14
    # column-wise backward error (graded columns keep their relative accuracy)
    err = np.linalg.norm(Q @ R - A, axis=0)
    assert np.all(err <= 1e-14 * np.maximum(np.linalg.norm(A, axis=0), 1e-300) * 30 + 1e-300)


@pytest.mark.parametrize("m,n,chi", [(6000, 96, 40), (2304, 576, 24), (40000, 64, 64)])
def test_tall_svd_qr_path_equals_gram_jacobi_path_and_lapack(tk, ctx, m, n, chi):
    """svd_trunc of a tall matrix: Householder QR + Jacobi of R (default) and the round-1
    Gram-preconditioned Jacobi ("disable_qr") give the LAPACK spectrum to 1e-12 sigma_1, and the
    QR path really ran."""
    rng = np.random.default_rng(m)
    A = rng.standard_normal((m, n)) @ np.diag(np.logspace(0, -9, n)) @ rng.standard_normal((n, n))
    sref = np.linalg.svd(A, compute_uv=False)
    out = {}
    for mode in (0, 1):
        ctx.set_option("disable_qr", mode)
        ctx.set_option("disable_subspace", 1)
        try:
            q0 = _counter(ctx, "qr_factorizations")
            U, S, Vt, eps = tk.svd_trunc(tk.DeviceTensor.from_numpy(A), 1, chi)
            ran = _counter(ctx, "qr_factorizations") - q0
        finally:
            ctx.set_option("disable_qr", 0)
            ctx.set_option("disable_subspace", 0)
        assert (ran >= 1) == (mode == 0)
        s = S.to_numpy()
        assert np.abs(s - sref[:chi]).max() <= 1e-12 * sref[0]
        u, vt = U.to_numpy(), Vt.to_numpy()
        rec = (u * s) @ vt
        best = np.linalg.svd(A, full_matrices=False)
        ref = (best[0][:, :chi] * best[1][:chi]) @ best[2][:chi]
        assert np.abs(rec - ref).max() <= 1e-11 * sref[0]
        assert abs(eps - np.linalg.norm(sref[chi:])) <= 1e-12 * sref[0]
        out[mode] = s
    assert np.abs(out[0] - out[1]).max() <= 1e-12 * sref[0]


@pytest.mark.parametrize("n", [48, 128, 256, 600])
def test_persistent_jacobi_equals_per_round_launches(tk, ctx, n):
    """The one-launch cooperative Jacobi (grid barrier per round, convergence test on the device)
    and the round-1 form (one launch per round, host sync per sweep) run the same rotations in
    the same order: identical spectra; both match LAPACK."""
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n + 7))
    M = A @ A.T          # positive semidefinite, as every matrix the schemes hand to eigh_trunc!
    res = {}
    for mode in (0, 1):
        ctx.set_option("disable_persistent_jacobi", mode)
        try:
            p0 = _counter(ctx, "persistent_jacobi")
            w, V, eps = tk.eigh_trunc(tk.DeviceTensor.from_numpy(M), n)
            used = _counter(ctx, "persistent_jacobi") - p0
        finally:
            ctx.set_option("disable_persistent_jacobi", 0)
        assert (used >= 1) == (mode == 0)
        res[mode] = (w.to_numpy(), V.to_numpy())
    wref = np.linalg.eigvalsh(M)
    wref = wref[np.argsort(-np.abs(wref))]
    for mode in (0, 1):
        w, V = res[mode]
        assert np.abs(w - wref).max() <= 1e-12 * np.abs(wref).max()
        assert np.abs(V.T @ V - np.eye(n)).max() <= 1e-12
        assert np.abs(M @ V - V * w).max() <= 1e-11 * np.abs(wref).max()
    assert np.array_equal(res[0][0], res[1][0])


def test_jacobi_sweep_limit_raises(tk, ctx):
    """A Jacobi iteration that does not converge within the sweep limit is an ERROR, not a silent
    result: provoked by a limit of one sweep, on both execution forms."""
    rng = np.random.default_rng(3)
    A = rng.standard_normal((200, 200))
    for mode in (0, 1):
        ctx.set_option("disable_persistent_jacobi", mode)
        ctx.set_option("jacobi_max_sweeps", 1)
        n0 = _counter(ctx, "jacobi_not_converged")
        try:
            with pytest.raises(tk.TNRCudaError, match="did not converge"):
                tk.svd_trunc(tk.DeviceTensor.from_numpy(A), 1, 10)
        finally:
            ctx.set_option("jacobi_max_sweeps", 40)
            ctx.set_option("disable_persistent_jacobi", 0)
        assert _counter(ctx, "jacobi_not_converged") == n0 + 1
    U, S, Vt, _ = tk.svd_trunc(tk.DeviceTensor.from_numpy(A), 1, 10)     # the engine still works
    assert np.abs(S.to_numpy() - np.linalg.svd(A, compute_uv=False)[:10]).max() <= 1e-12 * 30
