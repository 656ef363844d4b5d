#!/bin/bash
# Round 2, GPU call 1: whole device suite WITHOUT -x (every failure in one call), the time-box
# logic of bench.py at a small chi, and the permute variants with their ncu launch list.
#   gpurun --timeout 2400 -- 'bash tools/r02_call1.sh'
set -u
OUT=gpurun_out/r02_call1
mkdir -p "$OUT"
step() {  # step <name> <timeout-seconds> <command...>
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc  $(( $(date +%s) - t0 )) s  $(tail -n 1 "$OUT/$name.log" | cut -c1-200)" | tee -a "$OUT/summary.txt"
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > "$OUT/gpu.txt" 2>&1
step tests_all 1500 python -m pytest tests -q -m gpu --durations=25
step bench_timebox_chi12 300 python bench.py --chi 12 --steps 20 --warmup 5 --time-budget 100
step bench_chi8_all 200 python bench.py --chi 8 --steps 4 --warmup 4 --no-cpu-baseline
step permute_perf_24 300 python tools/permute_perf.py 24
step permute_ncu 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:copy_ -c 120 --csv --log-file "$OUT/permute_ncu.csv" python tools/permute_perf.py 24
cat "$OUT/summary.txt"
