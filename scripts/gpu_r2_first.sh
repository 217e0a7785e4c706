#!/bin/bash
# round 2, first GPU call: parity of the warp-mode engine + A/B against the cooperative products
mkdir -p gpurun_out
{
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for mode in warp coop; do
  if [ $mode = coop ]; then export BDRT_COOP=1; else unset BDRT_COOP; fi
  echo "== engine $mode"; timeout 120 python scripts/gpu_time_engine.py 2>&1 | grep "^model"
  echo "== map $mode"; timeout 300 python scripts/gpu_time_map.py 12500 50000 2>&1 | grep "^B=\|status"
  echo "== nuts $mode"; timeout 300 python scripts/gpu_time_nuts.py 1184 2 200 200 2>&1 | grep "^B=\|stepsize"
done
} > gpurun_out/r2_first.log 2>&1
cat gpurun_out/r2_first.log
