#!/bin/bash
# Build libpymfb.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -Xcompiler -Wall -shared ${PYMFB_NVCC_EXTRA} \
  -o ${PYMFB_OUT:-../libpymfb.so} pymfb.cu -ldl -lpthread
echo "built ${PYMFB_OUT:-$(cd .. && pwd)/libpymfb.so}"
