#!/bin/bash
mkdir -p gpurun_out
{
for mode in warp coop; do
  if [ $mode = coop ]; then export BDRT_COOP=1; else unset BDRT_COOP; fi
  for rep in 1 2; do
  echo "== map $mode"; timeout 300 python scripts/gpu_time_map.py 12500 50000 2>&1 | grep "^B=" | tail -1
  done
  echo "== map uniform $mode"; timeout 300 python scripts/gpu_time_map.py 4736 2000 2>&1 | grep "^B=" | tail -1
  echo "== nuts $mode"; timeout 300 python scripts/gpu_time_nuts.py 1184 2 200 200 2>&1 | grep "^B="
done
} > gpurun_out/r2_ab2.log 2>&1
cat gpurun_out/r2_ab2.log
