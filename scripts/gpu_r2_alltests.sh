#!/bin/bash
# the whole GPU suite + the Newton polish timing
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q 2>&1 | tail -25 | cut -c1-600 > gpurun_out/r2_alltests.log
cat gpurun_out/r2_alltests.log
timeout 300 python scripts/gpu_time_misc.py newton 2>&1 | tail -1 | tee gpurun_out/r2_newton_time.log
