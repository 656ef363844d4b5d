#!/bin/bash
# The driver's own test command on the final build, ATRG_3D chi=48 through bench.py at N=1, and the
# headline workload on the INT8 CRT engine as a second, clearly labelled bench line.
set -u
OUT=gpurun_out/r02_call17
mkdir -p "$OUT"
step() {
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2> "$OUT/$name.err"
  local rc=$?
  echo "rc=$rc  $(( $(date +%s) - t0 )) s  $(tail -n 1 "$OUT/$name.log" | cut -c1-400)" | tee -a "$OUT/summary.txt"
}
step pytest_gpu 900 python -m pytest tests -x -q -m gpu --durations=10
step smoke 120 python -c "import __graft_entry__ as g; g.smoke()"
step bench_atrg3d_chi48_n1 200 python bench.py --workload atrg3d --chi 48 --steps 3 --warmup 4 --time-budget 180
step bench_n1_chi24_ozaki_crt 420 python bench.py --engine ozaki_crt --steps 2 --warmup 2 --time-budget 380 --no-cpu-baseline
cat "$OUT/summary.txt"
