#!/bin/bash
# 8 GPUs of one box: the sharded headline at N = 8 (short budget), then in parallel on disjoint
# GPUs: ATRG_3D chi = 48 at N = 4 and N = 2, the 2-GPU parity tests incl. the beta sweep.
set -u
OUT=gpurun_out/r02_call11
mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpus.txt" 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 420 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --time-budget 170 \
  > "$OUT/bench_n8_chi24.log" 2> "$OUT/bench_n8_chi24.err"
echo "n8 rc=$?" | tee -a "$OUT/summary.txt"
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 500 $TR --nproc-per-node 4 --master-port 29542 bench.py --gpus 4 --workload atrg3d --chi 48 --rfactor gram --steps 3 --warmup 3 --time-budget 450 \
    > "$OUT/bench_atrg3d_chi48_n4.log" 2> "$OUT/bench_atrg3d_chi48_n4.err"; echo "atrg n4 rc=$?" >> "$OUT/summary.txt" ) &
( CUDA_VISIBLE_DEVICES=4,5 timeout 500 $TR --nproc-per-node 2 --master-port 29543 bench.py --gpus 2 --workload atrg3d --chi 48 --rfactor gram --steps 3 --warmup 3 --time-budget 450 \
    > "$OUT/bench_atrg3d_chi48_n2.log" 2> "$OUT/bench_atrg3d_chi48_n2.err"; echo "atrg n2 rc=$?" >> "$OUT/summary.txt" ) &
( CUDA_VISIBLE_DEVICES=6,7 timeout 500 python -m pytest tests/test_gpu_multi.py -q -m gpu > "$OUT/tests_multi.log" 2>&1; echo "tests rc=$?" >> "$OUT/summary.txt" ) &
wait
cat "$OUT/summary.txt"; tail -c 600 "$OUT/bench_n8_chi24.log"
