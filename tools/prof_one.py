"""One DMMA GEMM launch for ncu (TN, the layout of the HOTRG_3D chunk contraction)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tnrkit.jl_b200 as tk
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = tk.default_context()
A = torch.randn((n, n), dtype=torch.float64, device="cuda")
B = torch.randn((n, n), dtype=torch.float64, device="cuda")
C = torch.empty((n, n), dtype=torch.float64, device="cuda")
for _ in range(2):
    ctx.call("tnr_gemm", b"T", b"N", n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, C.data_ptr(), n)
torch.cuda.synchronize()
