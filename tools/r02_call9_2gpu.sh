#!/bin/bash
# 2 GPUs: every multi-GPU parity test + the sharded bench paths (small chi, then chi = 24 short)
set -u
OUT=gpurun_out/r02_call9
mkdir -p "$OUT"
step() {
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc  $(( $(date +%s) - t0 )) s  $(tail -n 1 "$OUT/$name.log" | cut -c1-400)" | tee -a "$OUT/summary.txt"
}
nvidia-smi -L > "$OUT/gpus.txt" 2>&1
step tests_multi 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_atrg3d_sym_sharded.py tests/test_gpu_atrg3d_factored.py -q -m gpu -k "two_gpus or sharded" --durations=5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
step bench_n2_chi12 300 $TR bench.py --gpus 2 --chi 12 --steps 4 --warmup 3
step bench_n1_chi12 300 python bench.py --gpus 1 --chi 12 --steps 4 --warmup 3 --no-cpu-baseline
step bench_atrg_n2_chi16 600 $TR bench.py --gpus 2 --workload atrg3d --chi 16 --steps 3 --warmup 3
step bench_n2_chi24 700 $TR bench.py --gpus 2 --chi 24 --steps 20 --warmup 5 --time-budget 400
cat "$OUT/summary.txt"
