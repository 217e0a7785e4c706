#!/bin/bash
# full GPU test suite + timing of the solvers
mkdir -p gpurun_out
{
echo "== tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "== engine"; timeout 120 python scripts/gpu_time_engine.py 2>&1 | grep "^model"
echo "== map"; timeout 300 python scripts/gpu_time_map.py 12500 50000 2>&1 | grep "^B=\|status\|lp mean"
echo "== nuts"; timeout 300 python scripts/gpu_time_nuts.py 1184 2 200 200 2>&1 | grep "^B=\|stepsize"
} > gpurun_out/r2_full.log 2>&1
cat gpurun_out/r2_full.log
