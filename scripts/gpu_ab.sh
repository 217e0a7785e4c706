#!/bin/bash
# A/B timing of library builds on one box (scratch_libs/*.so vs the in-tree build)
mkdir -p gpurun_out
{
for rep in 1 2; do
for lib in scratch_libs/libbdrt_*.so bayes_drt_b200/libbdrt.so; do
  echo "== $lib"
  BDRT_LIB=$PWD/$lib timeout 300 python scripts/gpu_time_map.py 9472 2000 2>&1 | grep "^B="
done
done
} > gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log
