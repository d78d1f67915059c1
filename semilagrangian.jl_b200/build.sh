#!/bin/bash
# Builds libslb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
mkdir -p "$HERE/lib"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -Xcompiler -O2 -ccbin /usr/bin/g++ ${SLB_NVCC_EXTRA} \
  -shared -o "$HERE/lib/libslb200.so" "$HERE/csrc/slb_api.cu" -lcudart
echo "built $HERE/lib/libslb200.so"
