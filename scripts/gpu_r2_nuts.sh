#!/bin/bash
mkdir -p gpurun_out
{
echo "== nuts"; timeout 300 python scripts/gpu_time_nuts.py 1184 2 200 200 2>&1 | grep "^B=\|stepsize"
echo "== tests"; timeout 1800 python -m pytest tests/test_gpu_nuts.py tests/test_gpu_nuts_configs.py tests/test_gpu_per_spectrum.py tests/test_gpu_series_parallel.py -x -q 2>&1 | tail -6
} > gpurun_out/r2_nuts.log 2>&1
cat gpurun_out/r2_nuts.log
