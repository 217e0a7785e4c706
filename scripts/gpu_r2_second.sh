#!/bin/bash
mkdir -p gpurun_out
{
echo "== new tests"; timeout 1200 python -m pytest tests/test_gpu_nuts_configs.py tests/test_gpu_map_benchmark.py -x -q -s 2>&1 | tail -25
echo "== phase clocks"; BDRT_LIB=$PWD/scratch_libs/libbdrt_clk.so timeout 300 python scripts/gpu_phase_clocks.py 2>&1 | tail -8
echo "== main lib timing"; timeout 120 python scripts/gpu_time_engine.py 2>&1 | grep "^model"
timeout 300 python scripts/gpu_time_map.py 12500 50000 2>&1 | grep "^B=\|status"
} > gpurun_out/r2_second.log 2>&1
cat gpurun_out/r2_second.log
