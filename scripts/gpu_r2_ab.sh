#!/bin/bash
# A/B of library builds in scratch_libs/ (names given as arguments) on the MAP benchmark and on uniform work
mkdir -p gpurun_out
{
for rep in 1 2; do
  for lib in "$@"; do
    echo "== $lib rep $rep"
    BDRT_LIB=$PWD/scratch_libs/libbdrt_$lib.so timeout 300 python scripts/gpu_time_map.py 12500 50000 2>&1 | grep "^B=" | tail -1
    BDRT_LIB=$PWD/scratch_libs/libbdrt_$lib.so timeout 300 python scripts/gpu_time_map.py 4736 2000 2>&1 | grep "^B=" | tail -1
  done
done
} > gpurun_out/r2_ab.log 2>&1
cat gpurun_out/r2_ab.log
