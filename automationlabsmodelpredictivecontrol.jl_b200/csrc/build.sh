#!/bin/bash
# Builds libmpcb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libmpcb200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3,-Wall -cudart static ${MPCB_NVCC_EXTRA:-}"
$NVCC $FLAGS -shared mpcb_api.cu nmpc_api.cu admm_stream.cu host_design.cpp -o $OUT
echo "built $(realpath $OUT)"
