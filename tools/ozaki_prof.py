"""One emulated 13824^3 GEMM for ncu (ozaki_tile_kernel): `python tools/ozaki_prof.py [n] [crt]`
(crt: the 16-modulus CRT variant instead of the 8 digit planes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tnrkit.jl_b200 as tk
n = int(sys.argv[1]) if len(sys.argv) > 1 else 13824
ctx = tk.default_context()
if len(sys.argv) > 2 and sys.argv[2] == "crt":
    ctx.set_option("ozaki_crt", 16)
else:
    ctx.set_option("ozaki", 8)
A = torch.randn((n, n), dtype=torch.float64, device="cuda")
B = torch.randn((n, n), dtype=torch.float64, device="cuda")
C = torch.empty((n, n), dtype=torch.float64, device="cuda")
for _ in range(1):
    ctx.call("tnr_gemm_ozaki", n, n, n, A.data_ptr(), n, B.data_ptr(), n, C.data_ptr(), n)
torch.cuda.synchronize()
