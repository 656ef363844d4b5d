"""One steady-state RG step of a named configuration between cudaProfilerStart / Stop, for ncu:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches_hotrg64.csv python tools/profile_step.py HOTRG 64 4
    python tools/launch_share.py gpurun_out/launches_hotrg64.csv > profiles/r02_share_hotrg64.md

Arguments: scheme chi warm_steps [model]   (model: ising | ising_z2 | potts_z3 | ising3d).
Without ncu it prints the CUDA-event time of the profiled step and the launch counters."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tnrkit.jl_b200 as tk  # noqa: E402

scheme, chi, warm = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
model = sys.argv[4] if len(sys.argv) > 4 and "=" not in sys.argv[4] else "ising"
opts = [a for a in sys.argv[4:] if "=" in a]          # engine options, e.g. disable_qr=1
T = {"ising": lambda: tk.classical_ising(tk.Trivial), "ising_z2": lambda: tk.classical_ising(),
     "potts_z3": lambda: tk.classical_potts(3),
     "ising3d": lambda: tk.classical_ising_3D(tk.Trivial)}[model]()
kw = {"shard": False} if scheme == "HOTRG_3D" else {}
s = getattr(tk, scheme)(T, **kw)
trunc = tk.truncrank(chi)
ctx = tk.default_context()
for o in opts:
    k, v = o.split("=")
    ctx.set_option(k, int(v))
s.finalize()
for _ in range(warm):
    s.step(trunc)
    s.finalize()
torch.cuda.synchronize()
ctx.reset_counters()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
s.step(trunc)
n = s.finalize()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
c = ctx.counters()
print(f"{scheme} {model} chi={chi}: step {warm + 1} took {e0.elapsed_time(e1) / 1e3:.3f} s "
      f"(norm {n:.12e}); launches {c['launches']}, GEMM launches {c['gemm_launches']}, "
      f"GEMM flop {c['gemm_flops']:.3e}, dims {getattr(s.T, 'dims', None)} options {opts}")
