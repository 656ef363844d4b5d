#!/bin/bash
# CholeskyQR2 between the Rayleigh-Ritz checks of the factored ATRG_3D subspace SVDs: device tests,
# parity of the factored paths, ATRG_3D chi=48 again, the new BTRG Z2 chi=128 fixture.
set -u
OUT=gpurun_out/r02_call14
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
step pytest_factored 500 python -m pytest tests/test_gpu_atrg3d_factored.py tests/test_gpu_api_behaviour.py -x -q --durations=5
step pytest_btrg128 300 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -k "BTRG_ising or ATRG_3D"
step atrg48_gram_phases 300 python tools/atrg3d_bench.py --chi 48 --steps 4 --rfactor gram --phases
step atrg48_gram 300 python tools/atrg3d_bench.py --chi 48 --steps 5 --rfactor gram
cat "$OUT/summary.txt"
