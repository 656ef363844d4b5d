#!/bin/bash
# New pivoted-Cholesky R factors (tnr_psd_factor): device tests, parity of the gram paths, and the
# phase split of ATRG_3D chi=48 with it.
set -u
OUT=gpurun_out/r02_call13
mkdir -p "$OUT"
step() {
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc  $(( $(date +%s) - t0 )) s  $(tail -n 1 "$OUT/$name.log" | cut -c1-300)" | tee -a "$OUT/summary.txt"
}
step pytest_psd 300 python -m pytest tests/test_gpu_psd_factor.py -x -q
step pytest_gram 400 python -m pytest tests/test_gpu_atrg3d_factored.py -x -q -k "full_size or oracle_on_device"
step atrg48_gram_phases 400 python tools/atrg3d_bench.py --chi 48 --steps 4 --rfactor gram --phases
step atrg48_gram 300 python tools/atrg3d_bench.py --chi 48 --steps 4 --rfactor gram
cat "$OUT/summary.txt"
