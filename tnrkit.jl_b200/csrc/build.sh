#!/bin/bash
# Builds libtnrcuda.so for sm_100a (cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/.obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
pids=()
for f in gemm_dmma gemm_tma gemm_ozaki permute elementwise jacobi qr pchol tensor_ops schemes api; do
  if [ ! -f "$HERE/.obj/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/.obj/$f.o" ] || \
     [ "$HERE/common.cuh" -nt "$HERE/.obj/$f.o" ] || [ "$HERE/tensor.hpp" -nt "$HERE/.obj/$f.o" ] || \
     [ "$HERE/schemes.hpp" -nt "$HERE/.obj/$f.o" ] || [ "$HERE/crt_math.cuh" -nt "$HERE/.obj/$f.o" ] || \
     [ "$HERE/permute_plan.cuh" -nt "$HERE/.obj/$f.o" ] || \
     [ "$HERE/crt_tables.inc" -nt "$HERE/.obj/$f.o" ] || [ "$HERE/../../include/tnrcuda.h" -nt "$HERE/.obj/$f.o" ]; then
    $NVCC $FLAGS -c "$HERE/$f.cu" -o "$HERE/.obj/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libtnrcuda.so" "$HERE"/.obj/*.o -Xcompiler -fPIC -cudart static
echo "built $OUT/libtnrcuda.so"
