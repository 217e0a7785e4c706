#!/bin/bash
# first GPU pass of a session: parity tests, micro timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 120 python scripts/gpu_peak.py > gpurun_out/peak.log 2>&1; cat gpurun_out/peak.log
timeout 200 python scripts/gpu_time_engine.py > gpurun_out/engine.log 2>&1; cat gpurun_out/engine.log
timeout 300 python scripts/gpu_time_map.py 4736 2000 > gpurun_out/map_2000.log 2>&1; cat gpurun_out/map_2000.log
timeout 600 python scripts/gpu_time_map.py 4736 50000 > gpurun_out/map_50000.log 2>&1; cat gpurun_out/map_50000.log
timeout 600 python scripts/gpu_time_nuts.py 592 2 200 200 > gpurun_out/nuts.log 2>&1; cat gpurun_out/nuts.log
