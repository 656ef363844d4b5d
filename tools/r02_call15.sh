#!/bin/bash
# CholeskyQR2 in the C++ subspace solvers + Jacobi block cap 8: targeted device tests, A/B timings of
# configs[1-2], ATRG_3D chi=48 (tsqr vs gram parity, device RNG, block sizes), CRT engine speed.
set -u
OUT=gpurun_out/r02_call15
mkdir -p "$OUT"
step() {
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a "$OUT/summary.txt"
  local t0=$(date +%s)
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc  $(( $(date +%s) - t0 )) s  $(tail -n 1 "$OUT/$name.log" | cut -c1-400)" | tee -a "$OUT/summary.txt"
}
step pytest_targeted 600 python -m pytest tests/test_gpu_psd_factor.py tests/test_gpu_primitives.py tests/test_gpu_schemes.py tests/test_gpu_baseline_sizes.py tests/test_gpu_symmetric.py tests/test_gpu_atrg3d_factored.py tests/test_gpu_reference_testsets.py -x -q --durations=8
OLD="jacobi_max_bc=16 disable_cholqr=1"
step t_hotrg64_new 200 python tools/profile_step.py HOTRG 64 4
step t_hotrg64_old 200 python tools/profile_step.py HOTRG 64 4 ising $OLD
step t_btrg128z2_new 200 python tools/profile_step.py BTRG 128 4 ising_z2
step t_btrg128z2_old 200 python tools/profile_step.py BTRG 128 4 ising_z2 $OLD
step t_potts_new 200 python tools/profile_step.py TRG 128 4 potts_z3
step t_potts_old 200 python tools/profile_step.py TRG 128 4 potts_z3 $OLD
step t_potts_bc16 200 python tools/profile_step.py TRG 128 4 potts_z3 jacobi_max_bc=16
step t_potts_bc4 200 python tools/profile_step.py TRG 128 4 potts_z3 jacobi_max_bc=4
step t_trg128_new 200 python tools/profile_step.py TRG 128 4
step t_atrg64_new 200 python tools/profile_step.py ATRG 64 4
step atrg48_gram 300 python tools/atrg3d_bench.py --chi 48 --steps 5 --rfactor gram
step atrg48_tsqr 400 python tools/atrg3d_bench.py --chi 48 --steps 4 --rfactor tsqr
step atrg48_gram_b160 300 python tools/atrg3d_bench.py --chi 48 --steps 4 --rfactor gram --block 160
step atrg48_gram_b80 300 python tools/atrg3d_bench.py --chi 48 --steps 4 --rfactor gram --block 80
step crt_speed 300 python tools/ozaki_crt_check.py --speed-only
cat "$OUT/summary.txt"
