#!/bin/bash
# 2 GPUs: multi-GPU parity tests (incl. the sharded gram path of the factored ATRG_3D step) and
# ATRG_3D chi=48 at N=2 through bench.py.
set -u
OUT=gpurun_out/r02_call16
mkdir -p "$OUT"
step() {
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2> "$OUT/$name.err"
  local rc=$?
  echo "rc=$rc  $(( $(date +%s) - t0 )) s  $(tail -n 1 "$OUT/$name.log" | cut -c1-400)" | tee -a "$OUT/summary.txt"
}
nvidia-smi -L > "$OUT/gpus.txt" 2>&1
step tests_multi 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_atrg3d_sym_sharded.py tests/test_gpu_atrg3d_factored.py -q -m gpu -k "two_gpus or sharded" --durations=5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
step bench_atrg3d_chi48_n2 300 $TR bench.py --gpus 2 --workload atrg3d --chi 48 --steps 3 --warmup 4 --time-budget 250
cat "$OUT/summary.txt"
