#!/bin/bash
# 2 GPUs, final build: ATRG_3D chi=48 through bench.py.
set -u
OUT=gpurun_out/r02_call21
mkdir -p "$OUT"
timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --workload atrg3d --chi 48 --steps 3 --warmup 4 --time-budget 70 > "$OUT/bench_atrg3d_chi48_n2.log" 2> "$OUT/bench_atrg3d_chi48_n2.err"
echo "rc=$? $(tail -n 1 "$OUT/bench_atrg3d_chi48_n2.log" | cut -c1-200)" | tee "$OUT/summary.txt"
