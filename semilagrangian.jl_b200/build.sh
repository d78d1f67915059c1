#!/bin/bash
# Builds libslb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
# Translation units are compiled in parallel, then linked.
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
mkdir -p "$HERE/lib" "$HERE/build"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 -ccbin /usr/bin/g++ ${SLB_NVCC_EXTRA}"
pids=()
for src in slb_api slb_pair slb_bspfused slb_bspsplit; do
  ( "$NVCC" $FLAGS -c -o "$HERE/build/$src.o" "$HERE/csrc/$src.cu" ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$HERE/lib/libslb200.so" "$HERE/build/slb_api.o" "$HERE/build/slb_pair.o" "$HERE/build/slb_bspfused.o" "$HERE/build/slb_bspsplit.o" -lcudart
echo "built $HERE/lib/libslb200.so"
