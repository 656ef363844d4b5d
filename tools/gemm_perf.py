"""Micro-benchmark of the DMMA GEMM against cuBLAS DGEMM (torch.matmul) on the same shapes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tnrkit.jl_b200 as tk

ctx = tk.default_context()
if os.environ.get("TNR_DISABLE_TMA"):
    ctx.set_option("disable_tma", 1)
    print("TMA kernel disabled (cp.async kernel for every layout)")
shapes = [("T", "N", 13824, 13824, 13824), ("T", "N", 6144, 6144, 13824), ("T", "N", 4096, 4096, 4096),
          ("T", "N", 576, 576, 331776), ("N", "N", 7962624 // 4, 576 // 4 * 4, 24), ("N", "N", 8192, 8192, 8192),
          ("N", "T", 8192, 8192, 8192), ("T", "T", 8192, 8192, 8192)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
for ta, tb, m, n, k in shapes:
    A = torch.randn((m, k) if ta == "T" else (k, m), dtype=torch.float64, device="cuda")  # col-major (k x m) / (m x k)
    B = torch.randn((n, k) if tb == "N" else (k, n), dtype=torch.float64, device="cuda")
    Cc = torch.empty((n, m), dtype=torch.float64, device="cuda")
    lda = k if ta == "T" else m
    ldb = k if tb == "N" else n
    def ours():
        ctx.call("tnr_gemm", ta.encode(), tb.encode(), m, n, k, 1.0, A.data_ptr(), lda, B.data_ptr(), ldb, 0.0, Cc.data_ptr(), m)
    # torch reference: row-major views.  col-major X(r x c) stored == torch tensor of shape (c, r)
    opA = A if ta == "T" else A.t()      # (m, k)
    opB = B.t() if tb == "N" else B      # (k, n)
    def ref():
        return torch.matmul(opA, opB)    # (m, n) row major
    ours(); r = ref(); torch.cuda.synchronize()
    err = (Cc.t() - r).abs().max().item() / max(1.0, r.abs().max().item())
    res = {}
    for name, fn in (("ours", ours), ("cublas", ref)):
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[name] = 2.0 * m * n * k / (best * 1e-3) / 1e12
    print(f"{ta}{tb} m={m} n={n} k={k}: ours {res['ours']:.2f} TF/s  cuBLAS {res['cublas']:.2f} TF/s  relerr {err:.1e}", flush=True)
    del A, B, Cc, r
    torch.cuda.empty_cache()
