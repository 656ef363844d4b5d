"""Opt-in INT8 Ozaki engine (tcgen05 + TMEM): FP64-level accuracy of the emulated GEMM and
parity of a HOTRG_3D run with the engine on against the FP64 tensor-core (DMMA) run."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _counter(ctx, name):
    v = C.c_double()
    ctx.call("tnr_get_counter", name.encode(), C.byref(v))
    return v.value


@pytest.mark.parametrize("m,n,k,wide", [(1024, 1536, 2048, False), (640, 768, 4096, True),
                                        (1000, 530, 1040, False)])
def test_gemm_ozaki_matches_extended_precision(tk, ctx, m, n, k, wide):
    rng = np.random.default_rng(m + n + k)
    A = rng.standard_normal((m, k))
    B = rng.standard_normal((n, k))
    if wide:  # eight orders of magnitude inside rows and across rows
        A *= np.exp(rng.uniform(-9, 9, size=(m, k)))
        B *= np.exp(rng.uniform(-9, 9, size=(n, 1)))
    dA, dB = tk.DeviceTensor.from_numpy(A.T), tk.DeviceTensor.from_numpy(B.T)  # K x M, K x N
    C1, C2 = tk.DeviceTensor.empty((m, n)), tk.DeviceTensor.empty((m, n))
    ctx.set_option("ozaki", 8)
    try:
        ctx.call("tnr_gemm_ozaki", m, n, k, dA.ptr, k, dB.ptr, k, C1.ptr, m)
    finally:
        ctx.set_option("ozaki", 0)
    ctx.call("tnr_gemm", b"T", b"N", m, n, k, 1.0, dA.ptr, k, dB.ptr, k, 0.0, C2.ptr, m)
    ref = A.astype(np.longdouble) @ B.T.astype(np.longdouble)
    scale = np.abs(A).astype(np.longdouble) @ np.abs(B.T).astype(np.longdouble)
    e_oz = float(np.max(np.abs(C1.to_numpy() - ref) / scale))
    e_dm = float(np.max(np.abs(C2.to_numpy() - ref) / scale))
    # componentwise-relative error (w.r.t. |A||B|) of the emulation is of the order of DGEMM's
    assert e_oz <= 4e-15, (e_oz, e_dm)
    assert e_dm <= 4e-15


def test_gemm_ozaki_refuses_unsupported_shapes(tk, ctx):
    a = tk.DeviceTensor.from_numpy(np.ones((600, 600)))
    c = tk.DeviceTensor.empty((600, 600))
    with pytest.raises(tk.TNRCudaError):      # engine disabled
        ctx.call("tnr_gemm_ozaki", 600, 600, 600, a.ptr, 600, a.ptr, 600, c.ptr, 600)
    ctx.set_option("ozaki", 8)
    try:
        with pytest.raises(tk.TNRCudaError):  # k not a multiple of 16
            ctx.call("tnr_gemm_ozaki", 600, 600, 600, a.ptr, 600, a.ptr, 600, c.ptr, 600)
    finally:
        ctx.set_option("ozaki", 0)


def test_hotrg3d_with_ozaki_engine_matches_dmma(tk, ctx):
    """chi = 12: the chunk contraction is 1728^3; with the engine on it runs as INT8 digit-plane
    products, and the norm list must agree with the FP64 tensor-core run to 1e-11."""
    T = tk.classical_ising_3D(tk.Trivial)
    ref = np.array(tk.run(tk.HOTRG_3D(T, shard=False), tk.truncrank(12), tk.maxiter(3), verbosity=0))
    before = _counter(ctx, "ozaki_gemms")
    ctx.set_option("ozaki", 8)
    try:
        got = np.array(tk.run(tk.HOTRG_3D(T, shard=False), tk.truncrank(12), tk.maxiter(3),
                              verbosity=0))
    finally:
        ctx.set_option("ozaki", 0)
    assert _counter(ctx, "ozaki_gemms") > before, "Ozaki engine was not used"
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= 1e-11
