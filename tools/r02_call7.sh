#!/bin/bash
set -u
OUT=gpurun_out/r02_call7
mkdir -p "$OUT"
step() {
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc  $(( $(date +%s) - t0 )) s  $(tail -n 1 "$OUT/$name.log" | cut -c1-300)" | tee -a "$OUT/summary.txt"
}
step tests_primitives 900 python -m pytest tests/test_gpu_primitives.py -q -m gpu
export TNR_TRACE=1
step trace_btrg128_z2 200 bash -c "python tools/profile_step.py BTRG 128 4 ising_z2 2>&1 | head -c 60000"
step trace_trg64 200 bash -c "python tools/profile_step.py TRG 64 4 2>&1 | head -c 30000"
unset TNR_TRACE
for cfg in "HOTRG 64 4" "TRG 64 4" "ATRG 64 4" "TRG 128 4 potts_z3"; do
  n=$(echo $cfg | tr ' ' '_')
  step time_${n}_qr 300 python tools/profile_step.py $cfg
  step time_${n}_noqr 300 python tools/profile_step.py $cfg disable_qr=1
  step time_${n}_qr_nopersist 300 python tools/profile_step.py $cfg disable_persistent_jacobi=1
done
cat "$OUT/summary.txt"
