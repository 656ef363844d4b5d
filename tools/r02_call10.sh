#!/bin/bash
# The driver's own commands on one GPU (headline + reference arm), then configs[3] at chi = 48.
set -u
OUT=gpurun_out/r02_call10
mkdir -p "$OUT"
step() {
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2> "$OUT/$name.err"
  local rc=$?
  echo "rc=$rc  $(( $(date +%s) - t0 )) s  $(tail -n 1 "$OUT/$name.log" | cut -c1-300)" | tee -a "$OUT/summary.txt"
}
step bench_n1_chi24 870 python bench.py --gpus 1 --steps 20 --warmup 5
step reference_arm 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1
step bench_atrg3d_chi48_gram 1500 python bench.py --workload atrg3d --chi 48 --rfactor gram --steps 3 --warmup 4 --time-budget 1300
cat "$OUT/summary.txt"
