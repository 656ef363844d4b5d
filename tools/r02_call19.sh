#!/bin/bash
# Final build: step times of configs[1-2], timings + ncu captures of the new Cholesky kernels.
set -u
OUT=gpurun_out/r02_call19
mkdir -p "$OUT"
step() {
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  echo "rc=$?  $(tail -n 1 "$OUT/$name.log" | cut -c1-300)" | tee -a "$OUT/summary.txt"
}
step pchol_time 60 python tools/pchol_prof.py
step t_hotrg64 60 python tools/profile_step.py HOTRG 64 4
step t_trg128 60 python tools/profile_step.py TRG 128 4
step t_btrg128z2 60 python tools/profile_step.py BTRG 128 4 ising_z2
step t_potts 60 python tools/profile_step.py TRG 128 4 potts_z3
step t_atrg64 60 python tools/profile_step.py ATRG 64 4
step ncu_pchol 90 ncu --clock-control none --set full --import-source on -k regex:"pchol_column|chol_inv" -s 600 -c 3 -f -o "$OUT/prof_pchol" python tools/pchol_prof.py
cat "$OUT/summary.txt"
