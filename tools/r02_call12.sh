#!/bin/bash
# Final 1-GPU validation with the shipped build: the driver's own test command, smoke(), the
# permute table, and the raw ncu evidence (full-set captures, launch lists).
set -u
OUT=gpurun_out/r02_call12
mkdir -p "$OUT"
step() {
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc  $(( $(date +%s) - t0 )) s  $(tail -n 1 "$OUT/$name.log" | cut -c1-300)" | tee -a "$OUT/summary.txt"
}
NCU="ncu --clock-control none"
step pytest_gpu 1200 python -m pytest tests -x -q -m gpu --durations=12
step smoke 120 python -c "import __graft_entry__ as g; g.smoke()"
step permute_perf_24 300 python tools/permute_perf.py 24
step permute_perf_32 300 python tools/permute_perf.py 32
step ncu_permute_rotate 300 $NCU --set full --import-source on -k regex:copy_bulk -s 3 -c 1 -f -o "$OUT/prof_copy_bulk_rotate" python tools/permute_one.py rotate
step ncu_permute_qnqk 300 $NCU --set full --import-source on -k regex:copy_bulk -s 3 -c 1 -f -o "$OUT/prof_copy_bulk_qnqk" python tools/permute_one.py qnqk
step launches_hotrg3d_chi16 900 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file "$OUT/launches_hotrg3d_chi16.csv" python tools/profile_step.py HOTRG_3D 16 3 ising3d
step share_hotrg64 600 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file "$OUT/launches_hotrg64.csv" python tools/profile_step.py HOTRG 64 4
step share_btrg128_z2 600 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file "$OUT/launches_btrg128_z2.csv" python tools/profile_step.py BTRG 128 4 ising_z2
step share_trg128_potts 600 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file "$OUT/launches_trg128_potts.csv" python tools/profile_step.py TRG 128 4 potts_z3
cat "$OUT/summary.txt"
