#!/bin/bash
# quick performance check of the solver kernels (no tests)
mkdir -p gpurun_out
T=${1:-perf}
{
timeout 200 python scripts/gpu_time_engine.py
timeout 300 python scripts/gpu_time_map.py 4736 2000
timeout 600 python scripts/gpu_time_map.py 4736 50000
timeout 600 python scripts/gpu_time_nuts.py 1184 2 200 200
} > gpurun_out/$T.log 2>&1
cat gpurun_out/$T.log
