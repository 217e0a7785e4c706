#!/bin/bash
# Build libbdrt.so in-tree for sm_100a (the only target).  Usage: build.sh [extra nvcc flags]
# Every .cu is compiled in parallel (nvcc -c), then linked; ptxas statistics of all files end up in build.log.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libbdrt.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v -static-global-template-stub=false"
mkdir -p build
pids=()
newest_hdr=$(ls -t *.cuh ../../include/*.h build.sh | head -1)
for f in *.cu; do
  o="build/${f%.cu}.o"
  # incremental: skip a file whose object is newer than its source and every header (BDRT_REBUILD=1 or extra flags force)
  if [ -z "$BDRT_REBUILD" ] && [ $# -eq 0 ] && [ -f "$o" ] && [ "$o" -nt "$f" ] && [ "$o" -nt "$newest_hdr" ]; then continue; fi
  ( $NVCC $FLAGS "$@" -c "$f" -o "build/${f%.cu}.o" > "build/${f%.cu}.log" 2>&1 ) &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
cat build/*.log > build.log
if [ $rc -ne 0 ]; then grep -E "error" -A3 build.log | head -60; exit 1; fi
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC build/*.o -o $OUT
grep -E "error|warning" build.log | grep -v "Wno" | head -20 || true
echo "built $(realpath $OUT)"
