#!/bin/bash
# Build an experimental variant of the library: scripts/build_variant.sh <name> <file.cu> <extra nvcc flags...>
# -> aria_b200/build/libaria_<name>.so (same C ABI; select with ARIA_B200_LIB=<path>).  Needs a prior `python -m aria_b200.build`.
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
B=aria_b200/build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -c aria_b200/csrc/$src -o $B/${src%.cu}_$name.o
objs=""
for o in gemm gemm2 gemm_wgrad moe_route moe_block moe_bwd ep elementwise attention attention_v3; do
  if [ "$o.cu" == "$src" ]; then objs="$objs $B/${o}_$name.o"; else objs="$objs $B/$o.o"; fi
done
nvcc -shared -o $B/libaria_$name.so $objs -gencode arch=compute_100a,code=sm_100a
echo $B/libaria_$name.so
