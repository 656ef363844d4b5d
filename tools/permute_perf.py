"""HBM throughput of the permute kernels on the shapes of a chi=24 HOTRG_3D step, for the default
kernels (permute_unroll = 1) and the opt-in variants with more loads in flight (2, 4).

    python tools/permute_perf.py [chi]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tnrkit.jl_b200 as tk
from tnrkit.jl_b200 import _lib
ctx = tk.default_context()
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 24
cases = [("rotate ((6,4),(2,3,1,5))", (chi,) * 6, (5, 3, 1, 2, 0, 4)),
         ("A2 -> [X|z b Y y x]", (chi,) * 6, (3, 0, 1, 2, 4, 5)),
         ("Pn -> Pk", (chi,) * 6, (0, 5, 4, 3, 1, 2)),
         ("Qn -> Qk", (chi,) * 6, (1, 3, 5, 0, 2, 4)),
         ("Gram operand [K|z x]", (chi,) * 6, (1, 2, 3, 4, 0, 5)),
         ("matrix transpose", (chi ** 3, chi ** 3), (1, 0)),
         ("tall transpose chi^4 x chi^2", (chi ** 4, chi ** 2), (1, 0)),
         ("ATRG_3D ((4,1),(2,3))-like", (chi,) * 6, (3, 0, 1, 2, 5, 4)),
         ("flat copy", (chi ** 6,), (0,))]
# (bulk, unroll, tile[, tpc, chunk_below, dense])
variants = [(1, 4, 96, 8, 0, 1), (1, 4, 96, 8, 0, 0), (1, 4, 96, 1, 0, 1), (0, 4, 96)] if len(sys.argv) < 3 else \
    [(0, 1, 96), (0, 2, 96), (0, 4, 96), (0, 4, 64), (0, 4, 48), (0, 1, 48), (0, 4, 32)]
for var in variants:
    bulk, unroll, tile = var[:3]
    tpc, chunk = (var[3], var[4]) if len(var) > 3 else (8, 0)
    dense = var[5] if len(var) > 5 else 1
    ctx.set_option("permute_dense", dense)
    ctx.set_option("permute_tpc", tpc)
    ctx.set_option("permute_chunk_below", chunk)
    ctx.set_option("permute_bulk", bulk)
    ctx.set_option("permute_unroll", unroll)
    ctx.set_option("permute_tile", tile)
    print(f"--- permute_bulk = {bulk}, permute_unroll = {unroll}, permute_tile = {tile}, "
          f"permute_tpc = {tpc}, permute_chunk_below = {chunk}, permute_dense = {dense}", flush=True)
    for name, dims, perm in cases:
        n = 1
        for d in dims: n *= d
        src = torch.randn(n, dtype=torch.float64, device="cuda")
        dst = torch.empty_like(src)
        def run():
            ctx.call("tnr_permute", src.data_ptr(), dst.data_ptr(), len(dims), _lib.i64(dims), _lib.i32(perm))
        run(); torch.cuda.synchronize()
        if (bulk, unroll, tile) != (0, 1, 96):   # same bits as the round-1 kernels
            ctx.set_option("permute_bulk", 0)
            ctx.set_option("permute_unroll", 1)
            ctx.set_option("permute_tile", 96)
            ref = torch.empty_like(src)
            ctx.call("tnr_permute", src.data_ptr(), ref.data_ptr(), len(dims), _lib.i64(dims), _lib.i32(perm))
            ctx.set_option("permute_bulk", bulk)
            ctx.set_option("permute_unroll", unroll)
            ctx.set_option("permute_tile", tile)
            torch.cuda.synchronize()
            assert torch.equal(ref, dst), name
            del ref
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{name:28s} {n*8/1e9:6.2f} GB  {best:8.3f} ms  {16.0*n/(best*1e-3)/1e9:8.1f} GB/s (read+write)", flush=True)
ctx.set_option("permute_dense", 1)
ctx.set_option("permute_tpc", 8)
ctx.set_option("permute_chunk_below", 0)
ctx.set_option("permute_bulk", 1)
ctx.set_option("permute_unroll", 4)
ctx.set_option("permute_tile", 96)
