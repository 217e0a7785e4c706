#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python scripts/gpu_time_map.py 12500 50000 2>&1 | tail -12
echo "== tests"; timeout 1200 python -m pytest tests/test_gpu_series_parallel.py tests/test_gpu_map.py tests/test_gpu_per_spectrum.py tests/test_gpu_inverter.py -x -q 2>&1 | tail -8
} > gpurun_out/r2_dbg.log 2>&1
cat gpurun_out/r2_dbg.log
