#!/bin/bash
# 4 GPUs: ATRG_3D chi=48 through bench.py (sharded factored step).
set -u
OUT=gpurun_out/r02_call18
mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpus.txt" 2>&1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --workload atrg3d --chi 48 --steps 3 --warmup 4 --time-budget 120 > "$OUT/bench_atrg3d_chi48_n4.log" 2> "$OUT/bench_atrg3d_chi48_n4.err"
echo "rc=$? $(tail -n 1 "$OUT/bench_atrg3d_chi48_n4.log" | cut -c1-300)" | tee "$OUT/summary.txt"
