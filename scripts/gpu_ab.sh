#!/bin/bash
# A/B of scratch_libs/libbdrt_prev.so against scratch_libs/libbdrt_new.so on one box (scripts/make_ab_libs.sh builds
# them): full MAP runs of the benchmark batch and / or the benchmark's HMC batch, two repetitions each, interleaved.
WHAT=${1:-both}
mkdir -p gpurun_out
{
for rep in 1 2; do
  for lib in prev new; do
    if [ "$WHAT" != nuts ]; then
      echo "== map $lib"; BDRT_LIB=$PWD/scratch_libs/libbdrt_$lib.so timeout 300 python scripts/gpu_time_map.py 12500 50000 2>&1 | grep "^B=" | tail -1
    fi
    if [ "$WHAT" != map ]; then
      echo "== nuts $lib"; BDRT_LIB=$PWD/scratch_libs/libbdrt_$lib.so timeout 300 python scripts/gpu_time_nuts.py 1184 2 200 200 2>&1 | grep "^B="
    fi
  done
done
} > gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log
