#!/bin/bash
# same-box A/B of two library builds (scratch_libs/libbdrt_prev.so vs scratch_libs/libbdrt_new.so) on the MAP driver
mkdir -p gpurun_out
{
for rep in 1 2; do
for lib in prev new; do
  export BDRT_LIB=$PWD/scratch_libs/libbdrt_$lib.so
  echo "== map $lib"; timeout 300 python scripts/gpu_time_map.py 12500 50000 2>&1 | grep "^B=" | tail -2
done
done
for lib in prev new; do
  export BDRT_LIB=$PWD/scratch_libs/libbdrt_$lib.so
  echo "== map uniform $lib"; timeout 300 python scripts/gpu_time_map.py 4736 2000 2>&1 | grep "^B=" | tail -1
done
} > gpurun_out/r2_ab3.log 2>&1
cat gpurun_out/r2_ab3.log
