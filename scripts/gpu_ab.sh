#!/bin/bash
# A/B timing of library builds on one box
mkdir -p gpurun_out
{
for lib in scratch_libs/libbdrt_*.so bayes_drt_b200/libbdrt.so; do
  echo "== $lib"
  BDRT_LIB=$PWD/$lib timeout 300 python scripts/gpu_time_map.py 9472 2000 2>&1 | grep "^B="
  BDRT_LIB=$PWD/$lib timeout 600 python scripts/gpu_time_nuts.py 1184 2 200 200 2>&1 | grep "^B="
done
} > gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log
