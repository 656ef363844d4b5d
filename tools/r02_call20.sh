#!/bin/bash
# Rayleigh-Ritz check placement from the previous RG step: device tests of the factored step and
# ATRG_3D chi=48 through bench.py.
set -u
OUT=gpurun_out/r02_call20
mkdir -p "$OUT"
timeout 120 python -m pytest tests/test_gpu_atrg3d_factored.py tests/test_gpu_psd_factor.py -x -q > "$OUT/pytest.log" 2>&1
echo "pytest rc=$? $(tail -n 1 "$OUT/pytest.log")" | tee "$OUT/summary.txt"
timeout 100 python bench.py --workload atrg3d --chi 48 --steps 3 --warmup 4 --time-budget 90 > "$OUT/bench_atrg3d_chi48_n1.log" 2> "$OUT/bench_atrg3d_chi48_n1.err"
echo "bench rc=$? $(tail -n 1 "$OUT/bench_atrg3d_chi48_n1.log" | cut -c1-200)" | tee -a "$OUT/summary.txt"
