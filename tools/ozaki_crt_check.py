"""Accuracy and speed of the CRT variant of the INT8 engine (`ozaki_crt` = 14..18 moduli) against
the digit-plane variant (`ozaki` = 8), the DMMA GEMM and an extended-precision reference.

    python tools/ozaki_crt_check.py [--speed-only]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tnrkit.jl_b200 as tk
ctx = tk.default_context()
rng = np.random.default_rng(0)


def engine(opt, val):
    ctx.set_option("ozaki", 0); ctx.set_option("ozaki_crt", 0)
    if opt: ctx.set_option(opt, val)


SPEED_ONLY = "--speed-only" in sys.argv     # accuracy is covered by tests/test_gpu_ozaki_crt.py
for (m, n, k, kind) in (() if SPEED_ONLY else ((1024, 1536, 2048, "randn"), (1024, 1024, 4096, "wide"))):
    A = rng.standard_normal((m, k)); B = rng.standard_normal((n, k))
    if kind == "wide":
        A *= np.exp(rng.uniform(-9, 9, size=(m, k))); B *= np.exp(rng.uniform(-9, 9, size=(n, 1)))
    dA = tk.DeviceTensor.from_numpy(A.T); dB = tk.DeviceTensor.from_numpy(B.T)
    ref = A.astype(np.longdouble) @ B.T.astype(np.longdouble)
    scale = np.abs(A).astype(np.longdouble) @ np.abs(B.T).astype(np.longdouble)
    for opt, val in (("ozaki_crt", 14), ("ozaki_crt", 15), ("ozaki_crt", 16), ("ozaki_crt", 17), ("ozaki", 8)):
        engine(opt, val)
        C1 = tk.DeviceTensor.empty((m, n))
        ctx.call("tnr_gemm_ozaki", m, n, k, dA.ptr, k, dB.ptr, k, C1.ptr, m)
        e = float(np.max(np.abs(C1.to_numpy() - ref) / scale))
        print(f"{opt}={val} {kind} {m}x{n}x{k}: max |err|/(|A||B|) = {e:.2e}", flush=True)
n = 13824
A = torch.randn((n, n), dtype=torch.float64, device="cuda"); B = torch.randn((n, n), dtype=torch.float64, device="cuda")
C = torch.empty((n, n), dtype=torch.float64, device="cuda")
for opt, val in (("ozaki_crt", 16), ("ozaki_crt", 15), ("ozaki", 8)):
    engine(opt, val)
    call = lambda: ctx.call("tnr_gemm_ozaki", n, n, n, A.data_ptr(), n, B.data_ptr(), n, C.data_ptr(), n)
    call(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{opt}={val} 13824^3 incl. splitting both operands and reconstruction: {best:.1f} ms -> "
          f"{2.0*n**3/(best*1e-3)/1e12:.1f} TFLOP/s FP64-equivalent", flush=True)
engine(None, 0)
