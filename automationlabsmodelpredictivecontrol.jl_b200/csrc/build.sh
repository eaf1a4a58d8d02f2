#!/bin/bash
# Builds libmpcb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).  The translation units are compiled in
# parallel (the templated kernels dominate the build time) and linked into one shared library with a static cudart.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${MPCB_OUT:-../libmpcb200.so}
OBJ=$(mktemp -d)
trap 'rm -rf "$OBJ"' EXIT
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3,-Wall ${MPCB_NVCC_EXTRA:-}"
pids=()
for src in mpcb_api.cu admm_smem.cu admm_smemg.cu admm_smemg_sig.cu admm_coop.cu nmpc_api.cu nmpc_variant_00.cu nmpc_variant_01.cu nmpc_variant_10.cu nmpc_variant_11.cu nmpc_variant_lin_00.cu nmpc_variant_lin_01.cu nmpc_variant_lin_10.cu nmpc_variant_lin_11.cu admm_stream.cu admm_riccati.cu host_design.cpp; do
  $NVCC $FLAGS -c "$src" -o "$OBJ/${src%.*}.o" > "$OBJ/${src%.*}.log" 2>&1 &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
cat "$OBJ"/*.log
[ $rc -eq 0 ] || { echo "compilation failed" >&2; exit 1; }
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -cudart static "$OBJ"/*.o -o $OUT
echo "built $(realpath $OUT)"
