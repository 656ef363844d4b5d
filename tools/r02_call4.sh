#!/bin/bash
set -u
OUT=gpurun_out/r02_call4
mkdir -p "$OUT"
step() {
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc  $(( $(date +%s) - t0 )) s  $(tail -n 1 "$OUT/$name.log" | cut -c1-200)" | tee -a "$OUT/summary.txt"
}
step tests 900 python -m pytest tests/test_gpu_permute_variants.py tests/test_gpu_schemes.py tests/test_gpu_primitives.py -q -m gpu
step permute_perf_24 300 python tools/permute_perf.py 24
step permute_perf_32 300 python tools/permute_perf.py 32
step permute_perf_16 300 python tools/permute_perf.py 16
step bench_chi16 300 python bench.py --chi 16 --steps 3 --warmup 3 --no-cpu-baseline
cat "$OUT/summary.txt"
