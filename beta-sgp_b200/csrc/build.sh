#!/bin/bash
# Build libbsgp.so in-tree for B200 (sm_100a).  Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
      -Xcompiler -fPIC -shared -o libbsgp.so bsgp_kernels.cu "$@"
