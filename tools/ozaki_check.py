"""Accuracy and speed of the INT8 Ozaki engine against the DMMA GEMM and float128-ish reference."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tnrkit.jl_b200 as tk
ctx = tk.default_context()
rng = np.random.default_rng(0)
for slices in (8,):
    ctx.set_option("ozaki", slices)
    for (m, n, k, kind) in ((1024, 1536, 2048, "randn"), (1024, 1024, 4096, "wide")):
        A = rng.standard_normal((m, k)); B = rng.standard_normal((n, k))
        if kind == "wide":  # dynamic range 1e8 inside rows and across rows
            A *= np.exp(rng.uniform(-9, 9, size=(m, k))); B *= np.exp(rng.uniform(-9, 9, size=(n, 1)))
        dA = tk.DeviceTensor.from_numpy(A.T)   # column-major K x M
        dB = tk.DeviceTensor.from_numpy(B.T)
        C1 = tk.DeviceTensor.empty((m, n)); C2 = tk.DeviceTensor.empty((m, n))
        ctx.call("tnr_gemm_ozaki", m, n, k, dA.ptr, k, dB.ptr, k, C1.ptr, m)
        ctx.call("tnr_gemm", b"T", b"N", m, n, k, 1.0, dA.ptr, k, dB.ptr, k, 0.0, C2.ptr, m)
        c1, c2 = C1.to_numpy(), C2.to_numpy()
        ref = (A.astype(np.longdouble) @ B.T.astype(np.longdouble))
        scale = (np.abs(A).astype(np.longdouble) @ np.abs(B.T).astype(np.longdouble))  # |A||B|
        e1 = float(np.max(np.abs(c1 - ref) / scale)); e2 = float(np.max(np.abs(c2 - ref) / scale))
        print(f"S={slices} {kind} {m}x{n}x{k}: max |err|/(|A||B|)  ozaki {e1:.2e}  dmma {e2:.2e}", flush=True)
# speed at the chunk size
n = 13824
A = torch.randn((n, n), dtype=torch.float64, device="cuda"); B = torch.randn((n, n), dtype=torch.float64, device="cuda")
C = torch.empty((n, n), dtype=torch.float64, device="cuda")
for slices in (7, 8):
    ctx.set_option("ozaki", slices)
    ctx.call("tnr_gemm_ozaki", n, n, n, A.data_ptr(), n, B.data_ptr(), n, C.data_ptr(), n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.call("tnr_gemm_ozaki", n, n, n, A.data_ptr(), n, B.data_ptr(), n, C.data_ptr(), n); e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"S={slices} 13824^3 incl. splitting both operands: {ms:.1f} ms -> {2.0*n**3/(ms*1e-3)/1e12:.1f} TFLOP/s FP64-equivalent", flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ctx.call("tnr_gemm", b"T", b"N", n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, C.data_ptr(), n)
e0.record(); ctx.call("tnr_gemm", b"T", b"N", n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, C.data_ptr(), n); e1.record(); e1.synchronize()
print(f"DMMA 13824^3: {e0.elapsed_time(e1):.1f} ms", flush=True)
