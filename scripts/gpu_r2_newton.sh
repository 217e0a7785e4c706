#!/bin/bash
# Newton polish after a change of newton.cu: its tests and its timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_map.py tests/test_gpu_map_benchmark.py -m gpu -q 2>&1 | tail -15 | cut -c1-400 > gpurun_out/r2_newton_tests.log
cat gpurun_out/r2_newton_tests.log
timeout 300 python scripts/gpu_time_misc.py newton 2>&1 | tail -1 | tee gpurun_out/r2_newton_time.log
