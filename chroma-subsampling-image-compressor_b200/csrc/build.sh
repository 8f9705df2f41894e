#!/usr/bin/env bash
# Builds libcsic.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libcsic.so
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC,-Wall,-fvisibility=hidden
       -Xptxas -v -cudart static)
"$NVCC" "${FLAGS[@]}" -o "$OUT" csic_params.cpp csic_api.cu csic_kernels.cu 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "ptxas info" || true
echo "built $(realpath $OUT)"
