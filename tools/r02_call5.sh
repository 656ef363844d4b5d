#!/bin/bash
set -u
OUT=gpurun_out/r02_call5
mkdir -p "$OUT"
step() {
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc  $(( $(date +%s) - t0 )) s  $(tail -n 1 "$OUT/$name.log" | cut -c1-200)" | tee -a "$OUT/summary.txt"
}
NCU="ncu --clock-control none"
step tests 1200 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_symmetric.py tests/test_gpu_permute_variants.py tests/test_gpu_baseline_sizes.py -q -m gpu --durations=10
step permute_perf_24 300 python tools/permute_perf.py 24
# --set full captures (one launch each) of the kernels the roofline numbers are about
step ncu_permute_qnqk 300 $NCU --set full --import-source on -k regex:copy_bulk -s 3 -c 1 -f -o "$OUT/prof_copy_bulk_qnqk" python tools/permute_one.py qnqk
step ncu_permute_rotate 300 $NCU --set full --import-source on -k regex:copy_bulk -s 3 -c 1 -f -o "$OUT/prof_copy_bulk_rotate" python tools/permute_one.py rotate
step ncu_gemm_13824 600 $NCU --set full --import-source on -k regex:gemm_dmma_tma -s 1 -c 1 -f -o "$OUT/prof_gemm_tma_13824" python tools/prof_one.py 13824
# launch-share lists of one steady step of the SVD-bound configurations (configs[1], configs[2])
step share_hotrg64 900 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file "$OUT/launches_hotrg64.csv" python tools/profile_step.py HOTRG 64 4
step share_trg128 900 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file "$OUT/launches_trg128.csv" python tools/profile_step.py TRG 128 4
step share_btrg128_z2 900 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file "$OUT/launches_btrg128_z2.csv" python tools/profile_step.py BTRG 128 4 ising_z2
step time_hotrg64 300 python tools/profile_step.py HOTRG 64 4
step time_trg128 300 python tools/profile_step.py TRG 128 4
step time_btrg128_z2 300 python tools/profile_step.py BTRG 128 4 ising_z2
cat "$OUT/summary.txt"
