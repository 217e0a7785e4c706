#!/bin/bash
# Build libbdrt.so in-tree for sm_100a (the only target).  Usage: build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRCS=$(ls *.cu)
OUT=../libbdrt.so
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
  -Xptxas -v "$@" $SRCS -o $OUT 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "Wno" | head -20 || true
echo "built $(realpath $OUT)"
