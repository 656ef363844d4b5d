#!/bin/bash
# Round 2, GPU call 2: the TMA-fed permute kernel (bit-exact tests, throughput, ncu DRAM bytes)
# and the time-box logic of bench.py with a budget that actually binds.
set -u
OUT=gpurun_out/r02_call2
mkdir -p "$OUT"
step() {
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc  $(( $(date +%s) - t0 )) s  $(tail -n 1 "$OUT/$name.log" | cut -c1-200)" | tee -a "$OUT/summary.txt"
}
step tests_permute 600 python -m pytest tests/test_gpu_permute_variants.py tests/test_gpu_primitives.py -q -m gpu
step permute_perf_24 300 python tools/permute_perf.py 24
step permute_perf_16 300 python tools/permute_perf.py 16
step permute_perf_48 300 python tools/permute_perf.py 32
step bench_timebox_chi16 300 python bench.py --chi 16 --steps 20 --warmup 5 --time-budget 45
step permute_ncu 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:copy_ -c 60 --csv --log-file "$OUT/permute_ncu.csv" python tools/permute_perf.py 24
cat "$OUT/summary.txt"
