"""CRT variant of the opt-in INT8 emulation engine (`tnr_set_option "ozaki_crt"` = 14..18 moduli:
ozaki_crt_split_kernel, ozaki_tile_kernel<true>, ozaki_crt_reconstruct_kernel in
csrc/gemm_ozaki.cu; scalar arithmetic in csrc/crt_math.cuh, host-checked by
tests/test_crt_math_host.py): accuracy against extended precision and against the bit-level model
(oracle/ozaki_model.py: multiply_crt), and a HOTRG_3D run with the engine on."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _counter(ctx, name):
    v = C.c_double()
    ctx.call("tnr_get_counter", name.encode(), C.byref(v))
    return v.value


@pytest.mark.parametrize("nmod,tol", [(16, 4e-16), (17, 4e-16), (15, 1e-14), (14, 3e-12)])
@pytest.mark.parametrize("m,n,k,wide", [(1024, 1536, 2048, False), (640, 768, 4096, True),
                                        (1000, 530, 1040, False)])
def test_gemm_ozaki_crt_matches_extended_precision(tk, ctx, m, n, k, wide, nmod, tol):
    rng = np.random.default_rng(m + n + k)
    A = rng.standard_normal((m, k))
    B = rng.standard_normal((n, k))
    if wide:
        A *= np.exp(rng.uniform(-9, 9, size=(m, k)))
        B *= np.exp(rng.uniform(-9, 9, size=(n, 1)))
    dA, dB = tk.DeviceTensor.from_numpy(A.T), tk.DeviceTensor.from_numpy(B.T)  # K x M, K x N
    C1 = tk.DeviceTensor.empty((m, n))
    ctx.set_option("ozaki_crt", nmod)
    try:
        ctx.call("tnr_gemm_ozaki", m, n, k, dA.ptr, k, dB.ptr, k, C1.ptr, m)
    finally:
        ctx.set_option("ozaki_crt", 0)
    got = C1.to_numpy()
    ref = A.astype(np.longdouble) @ B.T.astype(np.longdouble)
    scale = np.abs(A).astype(np.longdouble) @ np.abs(B.T).astype(np.longdouble)
    assert float(np.max(np.abs(got - ref) / scale)) <= tol
    # the bit-level model on a corner of the result: identical up to the final FP64 rounding
    import ozaki_model

    sub = ozaki_model.multiply_crt(A[:3], B[:4], nmod)
    assert np.max(np.abs(got[:3, :4] - sub) / np.abs(sub)) <= 4.5e-16


def test_hotrg3d_with_crt_engine_matches_dmma(tk, ctx):
    T = tk.classical_ising_3D(tk.Trivial)
    ref = np.array(tk.run(tk.HOTRG_3D(T, shard=False), tk.truncrank(12), tk.maxiter(3), verbosity=0))
    before = _counter(ctx, "ozaki_gemms")
    ctx.set_option("ozaki_crt", 16)
    try:
        got = np.array(tk.run(tk.HOTRG_3D(T, shard=False), tk.truncrank(12), tk.maxiter(3),
                              verbosity=0))
    finally:
        ctx.set_option("ozaki_crt", 0)
    assert _counter(ctx, "ozaki_gemms") > before, "CRT engine was not used"
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= 1e-11


def test_ozaki_crt_option_is_validated(tk, ctx):
    with pytest.raises(tk.TNRCudaError):
        ctx.set_option("ozaki_crt", 13)
    ctx.set_option("ozaki_crt", 0)
