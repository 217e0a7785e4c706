#!/bin/bash
mkdir -p gpurun_out
{
echo "== ubench"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/dmma_lat scripts/ubench/dmma_lat.cu && timeout 120 gpurun_out/dmma_lat; rm -f gpurun_out/dmma_lat
echo "== phase clocks"; BDRT_LIB=$PWD/scratch_libs/libbdrt_clk.so timeout 300 python scripts/gpu_phase_clocks.py 2>&1 | tail -8
echo "== map test"; timeout 1200 python -m pytest tests/test_gpu_map_benchmark.py -x -q -s 2>&1 | tail -12
} > gpurun_out/r2_third.log 2>&1
cat gpurun_out/r2_third.log
