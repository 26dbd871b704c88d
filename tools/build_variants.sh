#!/bin/bash
# Build tuning variants of librt_b200.so into build/variants/ (they travel to the GPU box).
# usage: tools/build_variants.sh "u16_b3:-DRT_SCAN_UNROLL=16 -DRT_TRACE_MIN_BLOCKS=3" ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/build/variants
mkdir -p $OUT
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -I$ROOT/include -c $ROOT/raytracingincuda_b200/csrc/rt_host.cpp -o $OUT/rt_host.o
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  (
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -I$ROOT/include $flags \
      -Xptxas -v -c $ROOT/raytracingincuda_b200/csrc/rt_kernels.cu -o $OUT/k_$name.o 2> $OUT/ptxas_$name.log
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/librt_b200_$name.so $OUT/k_$name.o $OUT/rt_host.o -cudart shared
  echo "$name: $(grep -A2 'trace_kernel_pbIfLi[03]' $OUT/ptxas_$name.log | grep -E 'Used' | head -1)"
  ) &
done
wait
