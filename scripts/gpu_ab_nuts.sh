#!/bin/bash
# A/B of two library builds on one box: NUTS throughput (1184 spectra x 2 chains x (200+200)) and a digest of the draws
mkdir -p gpurun_out
{
for lib in scratch_libs/libbdrt_prev.so bayes_drt_b200/libbdrt.so; do
  for rep in 1 2; do
    echo "== $lib"; BDRT_LIB=$PWD/$lib timeout 300 python scripts/gpu_time_nuts.py 1184 2 200 200 2>&1 | grep -E "^B=|stepsize"
  done
done
} > gpurun_out/ab_nuts.log 2>&1
cat gpurun_out/ab_nuts.log
