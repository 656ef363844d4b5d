#!/bin/bash
# First GPU call of round 2: everything that was written after round 1's GPU budget was spent.
#   gpurun --timeout 2400 -- 'bash tools/r02_first_gpu_call.sh'
# Every step has its own timeout and log under gpurun_out/r02_first/, so one failing or slow
# step does not cost the others.  Nothing here is a bench value (see bench.py for those).
set -u
OUT=gpurun_out/r02_first
mkdir -p "$OUT"
step() {  # step <name> <timeout-seconds> <command...>
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc  $(( $(date +%s) - t0 )) s  $(tail -n 1 "$OUT/$name.log" | cut -c1-160)" | tee -a "$OUT/summary.txt"
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > "$OUT/gpu.txt" 2>&1
# 1. the suite that was green in round 1 (regression check after the host-layer changes)
step tests_round1 900 python -m pytest tests -q -m gpu -x --durations=15 \
  --ignore=tests/test_gpu_models_u1.py --ignore=tests/test_gpu_atrg3d_factored.py \
  --ignore=tests/test_gpu_atrg3d_sym_sharded.py --ignore=tests/test_gpu_cft_observables.py \
  --ignore=tests/test_gpu_permute_variants.py --ignore=tests/test_gpu_ozaki_crt.py
# 2. device twins that have never run (no -x: collect every failure in one call)
step tests_models_u1 600 python -m pytest tests/test_gpu_models_u1.py -q -m gpu --durations=10
step tests_atrg3d_factored 900 python -m pytest tests/test_gpu_atrg3d_factored.py -q -m gpu --durations=10
step tests_atrg3d_sym_sharded 300 python -m pytest tests/test_gpu_atrg3d_sym_sharded.py -q -m gpu
step tests_cft 600 python -m pytest tests/test_gpu_cft_observables.py -q -m gpu --durations=10
step tests_permute_unroll 300 python -m pytest tests/test_gpu_permute_variants.py -q -m gpu
step tests_ozaki_crt 600 python -m pytest tests/test_gpu_ozaki_crt.py -q -m gpu
step ozaki_crt_check 600 python tools/ozaki_crt_check.py
# 3. permute variants (decides the default of permute_unroll / permute_tile)
step permute_perf_24 300 python tools/permute_perf.py 24
step permute_perf_16 120 python tools/permute_perf.py 16
# 3b. ncu evidence for the permute variants (durations and DRAM bytes per launch; not bench values)
step permute_ncu 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:copy_ -c 120 --csv --log-file "$OUT/permute_ncu.csv" python tools/permute_perf.py 24
# 4. configs[3] groundwork: ATRG_3D dense vs factored at chi=24, factored at chi=48
step atrg3d24_dense 300 python tools/atrg3d_bench.py --chi 24 --steps 4 --dense
step atrg3d24_factored 600 python tools/atrg3d_bench.py --chi 24 --steps 4 --phases
step atrg3d48_gram 1200 python tools/atrg3d_bench.py --chi 48 --steps 3 --rfactor gram --phases
cat "$OUT/summary.txt"
