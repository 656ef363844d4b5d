"""BASELINE.json configs[3]: ATRG_3D on classical_ising_3D(Trivial), seconds per RG step of the
factored step (tnrkit.jl_b200/atrg3d_factored.py), 1 GPU or sharded under torchrun:

    python tools/atrg3d_bench.py --chi 48 --steps 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 tools/atrg3d_bench.py --chi 48 --steps 3
    python tools/atrg3d_bench.py --chi 24 --steps 4 --dense      # tnr_atrg3d_step, for comparison

Per step: device time (CUDA events on the engine's stream, max over ranks), the subspace-iteration
counts and chunk plan of the last `_step!`, and the free energy after the run against the
reference's benchmark value (test/schemes.jl: f_benchmark3D = -3.507, rtol 5e-3 at chi = 12).
Not the headline bench (bench.py); numbers go to profiles/."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chi", type=int, default=48)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--dense", action="store_true", help="the dense tnr_atrg3d_step instead")
ap.add_argument("--block", type=int, default=None)
ap.add_argument("--rfactor", choices=("tsqr", "gram", "gram_eigh"), default="tsqr",
                help="R factors by chunked TSQR (2 chi^8 flop each) or from the Gram matrices of the "
                     "factors (chi^6)")
ap.add_argument("--phases", action="store_true",
                help="synchronise after every phase of a substep and report its seconds (adds "
                     "host syncs: use for the breakdown, not for the headline time)")
ap.add_argument("--max-chunk-elems", type=int, default=1 << 30)
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

import tnrkit.jl_b200 as tk  # noqa: E402
from tnrkit.jl_b200 import atrg3d_factored as af  # noqa: E402

af.PROFILE = bool(args.phases)
ctx = tk.default_context()
T = tk.classical_ising_3D(tk.Trivial)
if args.dense:
    s = tk.ATRG_3D(T, factored=False)
else:
    s = tk.ATRG_3D(T, factored=True, shard=world > 1, block=args.block,
                   max_chunk_elems=args.max_chunk_elems, rfactor=args.rfactor)
trunc = tk.truncrank(args.chi)
data = [s.finalize()]
rows = []
for it in range(args.steps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.reset_counters()
    e0.record()
    s.step(trunc)
    data.append(s.finalize())
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    c = ctx.counters()
    rows.append({"step": it + 1, "s": round(ms.item() / 1e3, 3),
                 "gemm_tflop": round(c["gemm_flops"] / 1e12, 3), "launches": c["launches"],
                 "stats": None if args.dense else json.loads(json.dumps(af.LAST_STATS))})
f = tk.free_energy(data, tk.ising_βc_3D, scalefactor=8.0)
if rank == 0:
    dims = list(s.factors.dims) if s.factors is not None else list(s.T.dims)
    print(json.dumps({"config": "ATRG_3D classical_ising_3D(Trivial)", "chi": args.chi,
                      "path": "dense" if args.dense else f"factored/{args.rfactor}", "n_gpus": world,
                      "dims": dims, "steps": rows,
                      "steady_s_per_step": min(r["s"] for r in rows[-2:]),
                      "free_energy": f, "rel_err_vs_f_benchmark3D": abs((f + 3.507) / 3.507),
                      "norms": data}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
