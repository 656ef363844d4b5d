"""The numpy stand-ins of the round-2 Cholesky entry points (tests/abi_emulator.py) against their
contracts in include/tnrcuda.h -- they carry the CPU tests of the factored ATRG_3D step, and the
device kernels are compared with them in tests/test_gpu_psd_factor.py."""
import ctypes as C

import numpy as np
import pytest

from abi_emulator import EmulatedContext

EPS = np.finfo(float).eps


def _factor(G):
    n = G.shape[0]
    g = np.asfortranarray(G)
    out = np.zeros((n, n), order="F")
    r = C.c_int64(-1)
    EmulatedContext()._tnr_psd_factor(g.ctypes.data, n, out.ctypes.data, C.byref(r))
    return out, int(r.value)


@pytest.mark.parametrize("n,rank", [(1, 1), (40, 40), (65, 20), (300, 120)])
def test_psd_factor_model(n, rank):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((3 * rank, rank)) @ rng.standard_normal((rank, n))
    G = A.T @ A
    L, r = _factor(G)
    assert r == rank and np.all(L[:, r:] == 0.0)
    assert np.abs(L @ L.T - G).max() <= 64 * n * EPS * np.diag(G).max()
    # what atrg3d.jl:58-66 uses: the singular values of R1 R2 are those of L1^T L2
    B = rng.standard_normal((n, 2 * n))
    L2, r2 = _factor(B @ B.T)
    s = np.linalg.svd(L[:, :r].T @ L2[:, :r2], compute_uv=False)
    ref = np.linalg.svd(np.linalg.qr(A, mode="r") @ np.linalg.qr(B.T, mode="r").T, compute_uv=False)
    k = min(len(s), len(ref), rank)
    assert np.abs(s[:k] - ref[:k]).max() <= 1e-11 * ref[0]


def test_psd_factor_model_zero_matrix():
    L, r = _factor(np.zeros((7, 7)))
    assert r == 0 and not L.any()


def test_orthonormalize_model():
    rng = np.random.default_rng(1)
    A = np.asfortranarray(rng.standard_normal((500, 30)) * np.logspace(0, -3, 30))
    Q = A.copy(order="F")
    refused = C.c_int(-1)
    EmulatedContext()._tnr_orthonormalize(Q.ctypes.data, 500, 30, C.byref(refused))
    assert refused.value == 0 and np.abs(Q.T @ Q - np.eye(30)).max() <= 1e-13
    R = Q.T @ A
    assert np.abs(np.tril(R, -1)).max() <= 1e-12 * np.abs(R).max() and np.all(np.diag(R) > 0)
    bad = A.copy(order="F")
    bad[:, 4] = bad[:, 2]
    keep = bad.copy()
    EmulatedContext()._tnr_orthonormalize(bad.ctypes.data, 500, 30, C.byref(refused))
    assert refused.value == 1 and np.array_equal(bad, keep)
    wide = np.asfortranarray(rng.standard_normal((400, 153)))
    EmulatedContext()._tnr_orthonormalize(wide.ctypes.data, 400, 153, C.byref(refused))
    assert refused.value == 1


def test_fill_random_model_is_deterministic_and_uniform():
    x, y = np.zeros(50001), np.zeros(50001)
    EmulatedContext()._tnr_fill_random(x.ctypes.data, x.size, 0x5EED)
    EmulatedContext()._tnr_fill_random(y.ctypes.data, y.size, 0x5EED)
    assert np.array_equal(x, y) and np.abs(x).max() < 1.0
    assert abs(x.mean()) < 0.02 and abs(x.std() - 1 / np.sqrt(3)) < 0.02
    EmulatedContext()._tnr_fill_random(y.ctypes.data, y.size, 0x5EEE)
    assert not np.array_equal(x, y)
