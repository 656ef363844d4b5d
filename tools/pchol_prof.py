"""Times tnr_psd_factor (n = 2304, the Gram matrices of ATRG_3D chi = 48) and tnr_orthonormalize
(110592 x 112, the subspace block of the same step) with CUDA events; under ncu the same script
gives the per-kernel captures (`-k regex:pchol_column|chol_inv`)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tnrkit.jl_b200 as tk  # noqa: E402

ctx = tk.default_context()
n = 2304
A = torch.randn(n, 3 * n, dtype=torch.float64, device="cuda") * torch.logspace(0, -6, n, dtype=torch.float64, device="cuda")[:, None]
G = (A @ A.T).contiguous()
L = torch.empty(n * n, dtype=torch.float64, device="cuda")
r = C.c_int64(0)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


ms = timed(lambda: ctx.call("tnr_psd_factor", G.data_ptr(), n, L.data_ptr(), C.byref(r)))
Lm = L.view(n, n).T          # column major -> torch row major transposed view
err = ((Lm @ Lm.T) - G).abs().max().item() / G.diagonal().max().item()
print(f"tnr_psd_factor n={n}: {ms:.2f} ms, rank {r.value}, max |L L^T - G| / max G_ii = {err:.2e}")
m, b = 110592, 112
Z = torch.randn(b, m, dtype=torch.float64, device="cuda")      # column-major m x b
Z0 = Z.clone()
ref = C.c_int(1)


def orth():
    Z.copy_(Z0)
    ctx.call("tnr_orthonormalize", Z.data_ptr(), m, b, C.byref(ref))


ms = timed(orth)
Q = Z                                                           # rows of Z are the columns of Q
dev = ((Q @ Q.T) - torch.eye(b, dtype=torch.float64, device="cuda")).abs().max().item()
print(f"tnr_orthonormalize {m} x {b}: {ms:.2f} ms incl. the restoring copy, refused {ref.value}, "
      f"max |Q^T Q - I| = {dev:.2e}")
