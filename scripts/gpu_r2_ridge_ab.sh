#!/bin/bash
mkdir -p gpurun_out
{
for lib in r256 r512; do echo "== $lib"; BDRT_LIB=$PWD/scratch_libs/libbdrt_$lib.so timeout 300 python scripts/gpu_time_ridge.py 2>&1 | tail -4; done
echo "== clocks (r256)"; BDRT_LIB=$PWD/scratch_libs/libbdrt_clk.so python scripts/gpu_ridge_clocks.py 2>&1 | tail -2
echo "== tests"; timeout 1200 python -m pytest tests/test_gpu_ridge.py tests/test_gpu_map_benchmark.py tests/test_gpu_summaries.py tests/test_gpu_inverter.py -q -x 2>&1 | tail -8 | cut -c1-300
} > gpurun_out/r2_ridge_ab.log 2>&1
cat gpurun_out/r2_ridge_ab.log
