#!/bin/bash
set -u
OUT=gpurun_out/r02_call6
mkdir -p "$OUT"
step() {
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc  $(( $(date +%s) - t0 )) s  $(tail -n 1 "$OUT/$name.log" | cut -c1-300)" | tee -a "$OUT/summary.txt"
}
step tests_qr 600 python -m pytest tests/test_gpu_primitives.py -q -m gpu -x -k "qr or svd or eigh or jacobi or orth"
step tests_schemes 900 python -m pytest tests/test_gpu_schemes.py tests/test_gpu_symmetric.py tests/test_gpu_atrg3d_factored.py tests/test_gpu_permute_variants.py -q -m gpu --durations=8
step permute_perf_24 300 python tools/permute_perf.py 24
for cfg in "HOTRG 64 4" "TRG 128 4" "BTRG 128 4 ising_z2" "TRG 128 4 potts_z3" "ATRG 64 4"; do
  n=$(echo $cfg | tr ' ' '_')
  step time_${n}_qr 300 python tools/profile_step.py $cfg
  step time_${n}_noqr 300 python tools/profile_step.py $cfg disable_qr=1
done
step atrg3d24_dense 300 python tools/atrg3d_bench.py --chi 24 --steps 4 --dense
step atrg3d24_factored 600 python tools/atrg3d_bench.py --chi 24 --steps 5
step bench_atrg3d_16 300 python bench.py --workload atrg3d --chi 16 --steps 3 --warmup 3
step atrg3d32_gram 900 python tools/atrg3d_bench.py --chi 32 --steps 5 --rfactor gram --phases
step atrg3d32_tsqr 900 python tools/atrg3d_bench.py --chi 32 --steps 5 --phases
cat "$OUT/summary.txt"
