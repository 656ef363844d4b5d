"""tnr_psd_factor (csrc/pchol.cu): the rank-revealing factor L L^T = G of a symmetric positive
semidefinite matrix that gives the factored ATRG_3D step its R factors (atrg3d.jl:53-56, up to
the left orthogonal gauge) -- against numpy, and against the column-by-column restatement of the
same algorithm in tests/abi_emulator.py (same pivots, same clamps)."""
import ctypes as C

import numpy as np
import pytest

from abi_emulator import EmulatedContext

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def _psd(rng, n, rank, decay):
    A = rng.standard_normal((max(rank, 1) * 3, n))
    if rank < n:
        A = rng.standard_normal((rank * 3, rank)) @ rng.standard_normal((rank, n))
    A *= decay ** (np.arange(n) / max(1, n - 1) * 16)            # graded columns: 16 decades at decay = 0.1
    return A.T @ A


def _factor(tk, ctx, G):
    n = G.shape[0]
    Gd = tk.DeviceTensor.from_numpy(G)
    L = tk.DeviceTensor.empty((n, n), 1, ctx)
    r = C.c_int64(-7)
    ctx.call("tnr_psd_factor", Gd.ptr, n, L.ptr, C.byref(r))
    assert np.array_equal(Gd.to_numpy(), G)                      # input not modified
    return L.to_numpy(), int(r.value)


def _model(G):
    n = G.shape[0]
    g = np.asfortranarray(G)
    out = np.zeros((n, n), order="F")
    r = C.c_int64(0)
    EmulatedContext()._tnr_psd_factor(g.ctypes.data, n, out.ctypes.data, C.byref(r))
    return out, int(r.value)


@pytest.mark.parametrize("n,rank,decay", [(1, 1, 1.0), (7, 7, 1.0), (33, 33, 1.0), (64, 64, 1.0),
                                          (65, 20, 1.0), (200, 200, 0.1), (257, 100, 0.5),
                                          (1000, 1000, 1.0), (1024, 300, 0.1), (2304, 2304, 0.1)])
def test_psd_factor_reproduces_the_matrix(tk, ctx, n, rank, decay):
    rng = np.random.default_rng(n * 31 + rank)
    G = _psd(rng, n, rank, decay)
    L, r = _factor(tk, ctx, G)
    scale = np.abs(np.diag(G)).max()
    assert np.abs(L @ L.T - G).max() <= 64 * n * EPS * scale
    assert 1 <= r <= n and np.all(L[:, r:] == 0.0)
    if rank < n:
        assert r <= rank + 2                                      # numerical rank found
    if decay == 1.0 and rank == n:
        assert r == n
    # the CPU restatement of the algorithm takes the same pivots: same factor up to rounding
    Lm, rm = _model(G)
    assert np.abs(Lm @ Lm.T - G).max() <= 64 * n * EPS * scale
    k = min(r, rm, 8)
    assert np.abs(L[:, :k] - Lm[:, :k]).max() <= 1e-9 * np.sqrt(scale)


def test_psd_factor_zero_and_nonfinite(tk, ctx):
    L, r = _factor(tk, ctx, np.zeros((40, 40)))
    assert r == 0 and not L.any()
    G = np.eye(5)
    G[2, 2] = np.nan
    with pytest.raises(tk.TNRCudaError):
        _factor(tk, ctx, G)


def test_psd_factor_is_an_r_factor_for_the_projectors(tk, ctx):
    """What atrg3d.jl:58-66 needs: the singular values of R1 R2 from the Cholesky factors of the
    Gram matrices equal those from Householder R factors."""
    rng = np.random.default_rng(5)
    A = rng.standard_normal((4000, 144)) * np.logspace(0, -6, 144)
    B = rng.standard_normal((144, 3000)) * np.logspace(0, -6, 144)[:, None]
    L1, r1 = _factor(tk, ctx, A.T @ A)
    L2, r2 = _factor(tk, ctx, B @ B.T)
    s = np.linalg.svd(L1[:, :r1].T @ L2[:, :r2], compute_uv=False)
    R1 = np.linalg.qr(A, mode="r")
    R2 = np.linalg.qr(B.T, mode="r").T
    ref = np.linalg.svd(R1 @ R2, compute_uv=False)
    assert np.abs(s[:12] - ref[:12]).max() <= 1e-12 * ref[0]


def _orth(tk, ctx, A):
    m, n = A.shape
    Ad = tk.DeviceTensor.from_numpy(A)
    refused = C.c_int(-1)
    ctx.call("tnr_orthonormalize", Ad.ptr, m, n, C.byref(refused))
    return Ad.to_numpy(), refused.value


@pytest.mark.parametrize("m,n,cond", [(64, 1, 1.0), (300, 7, 10.0), (5000, 112, 1e3), (4096, 152, 1e2),
                                      (20000, 128, 1e3), (110592, 112, 1e4)])
def test_orthonormalize_cholqr2(tk, ctx, m, n, cond):
    """tnr_orthonormalize: Q^T Q = I to rounding, Q = A R^-1 with R upper triangular, positive
    diagonal (the thin QR factor with that sign convention)."""
    rng = np.random.default_rng(m + n)
    u, _ = np.linalg.qr(rng.standard_normal((m, n)))
    v, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = (u * np.logspace(0, -np.log10(cond), n)) @ v.T
    Q, refused = _orth(tk, ctx, A)
    assert refused == 0
    assert np.abs(Q.T @ Q - np.eye(n)).max() <= 5e-14
    R = Q.T @ A
    assert np.abs(np.tril(R, -1)).max() <= 1e-12 * np.abs(R).max() and np.all(np.diag(R) > 0)
    assert np.abs(Q @ R - A).max() <= 1e-12 * np.abs(A).max()
    before = ctx.counters()["launches"]
    _orth(tk, ctx, A)
    assert ctx.counters()["launches"] - before <= 16          # GEMMs, split-K reductions, 2 one-CTA kernels


def test_orthonormalize_refuses_ill_conditioned_and_leaves_input(tk, ctx):
    rng = np.random.default_rng(3)
    A = rng.standard_normal((2000, 40))
    A[:, 7] = A[:, 3] + 1e-9 * A[:, 5]                            # cond ~ 1e9
    Q, refused = _orth(tk, ctx, A)
    assert refused == 1 and np.array_equal(Q, A)
    A[:, 7] = A[:, 3]                                             # exactly rank deficient
    Q, refused = _orth(tk, ctx, A)
    assert refused == 1 and np.array_equal(Q, A)
    Z = np.zeros((100, 5))
    Q, refused = _orth(tk, ctx, Z)
    assert refused == 1
    W = rng.standard_normal((4096, 153))                          # more columns than fit in shared memory
    Q, refused = _orth(tk, ctx, W)
    assert refused == 1 and np.array_equal(Q, W)
    with pytest.raises(tk.TNRCudaError):
        _orth(tk, ctx, rng.standard_normal((5, 9)))               # wide


def test_fill_random_matches_the_host_restatement(tk, ctx):
    n = 100003
    x = tk.DeviceTensor.empty((n,), 1, ctx)
    ctx.call("tnr_fill_random", x.ptr, n, 0x5EED)
    ref = np.zeros(n)
    EmulatedContext()._tnr_fill_random(ref.ctypes.data, n, 0x5EED)
    got = x.to_numpy()
    assert np.array_equal(got, ref)
    assert abs(got.mean()) < 0.01 and abs(got.std() - 1 / np.sqrt(3)) < 0.01 and np.abs(got).max() < 1.0
