#!/bin/bash
mkdir -p gpurun_out
{
echo "== ridge timing"; timeout 300 python scripts/gpu_time_ridge.py 2>&1 | tail -4
echo "== map benchmark test"; timeout 1200 python -m pytest tests/test_gpu_map_benchmark.py -q -x -s 2>&1 | grep -v Deprecat | tail -25 | cut -c1-600
echo "== tests"; timeout 2400 python -m pytest tests/test_gpu_ridge.py tests/test_gpu_summaries.py tests/test_gpu_inverter.py -q 2>&1 | tail -12 | cut -c1-300
} > gpurun_out/r2_rest.log 2>&1
cat gpurun_out/r2_rest.log
