#!/bin/bash
# Builds libslb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
# Translation units are compiled in parallel (the pair-fused kernel once per stencil width), then linked.
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
mkdir -p "$HERE/lib" "$HERE/build"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 -ccbin /usr/bin/g++ ${SLB_NVCC_EXTRA}"
pids=()
objs=()
for src in slb_api slb_pair slb_bspfused slb_bspsplit slb_bspseg slb_comm slb_program; do
  [ -f "$HERE/csrc/$src.cu" ] || continue
  ( "$NVCC" $FLAGS -c -o "$HERE/build/$src.o" "$HERE/csrc/$src.cu" ) &
  pids+=($!)
  objs+=("$HERE/build/$src.o")
done
for p1 in 4 6 8 10 12; do
  ( "$NVCC" $FLAGS -DSLB_PAIR_P1=$p1 -c -o "$HERE/build/slb_pair_p$p1.o" "$HERE/csrc/slb_pair.cu" ) &
  pids+=($!)
  objs+=("$HERE/build/slb_pair_p$p1.o")
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$HERE/lib/libslb200.so" "${objs[@]}" -lcudart
echo "built $HERE/lib/libslb200.so"
