"""One permute case for ncu (copy_bulk_kernel):  python tools/permute_one.py <case> [chi]
cases: rotate | a2 | pnpk | qnqk | gram | transpose"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tnrkit.jl_b200 as tk  # noqa: E402
from tnrkit.jl_b200 import _lib  # noqa: E402

case = sys.argv[1]
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 24
CASES = {"rotate": ((chi,) * 6, (5, 3, 1, 2, 0, 4)), "a2": ((chi,) * 6, (3, 0, 1, 2, 4, 5)),
         "pnpk": ((chi,) * 6, (0, 5, 4, 3, 1, 2)), "qnqk": ((chi,) * 6, (1, 3, 5, 0, 2, 4)),
         "gram": ((chi,) * 6, (1, 2, 3, 4, 0, 5)), "transpose": ((chi ** 3, chi ** 3), (1, 0))}
dims, perm = CASES[case]
ctx = tk.default_context()
n = 1
for d in dims:
    n *= d
src = torch.randn(n, dtype=torch.float64, device="cuda")
dst = torch.empty_like(src)
for _ in range(3):
    ctx.call("tnr_permute", src.data_ptr(), dst.data_ptr(), len(dims), _lib.i64(dims), _lib.i32(perm))
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.call("tnr_permute", src.data_ptr(), dst.data_ptr(), len(dims), _lib.i64(dims), _lib.i32(perm))
    e1.record()
    e1.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"{case} chi={chi}: {best:.3f} ms  {16.0 * n / (best * 1e-3) / 1e9:.1f} GB/s (read+write)")
