#!/bin/bash
# quick check after an engine / solver change: parity tests, engine / MAP / NUTS timing
mkdir -p gpurun_out
{
echo "== tests"; timeout 1200 python -m pytest tests/test_gpu_logpost.py tests/test_gpu_series_parallel.py tests/test_gpu_map.py tests/test_gpu_per_spectrum.py tests/test_gpu_nuts.py -x -q 2>&1 | tail -8
echo "== engine"; timeout 120 python scripts/gpu_time_engine.py 2>&1 | grep "^model"
echo "== map"; timeout 300 python scripts/gpu_time_map.py 12500 50000 2>&1 | grep "^B=\|status\|lp mean"
echo "== map uniform"; timeout 300 python scripts/gpu_time_map.py 4736 2000 2>&1 | grep "^B=" | tail -1
echo "== nuts"; timeout 300 python scripts/gpu_time_nuts.py 1184 2 200 200 2>&1 | grep "^B=\|stepsize"
} > gpurun_out/r2_quick.log 2>&1
cat gpurun_out/r2_quick.log
