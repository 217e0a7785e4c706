#!/bin/bash
# ncu captures of the warp-mode kernels (one launch each, --set full + source counters)
mkdir -p gpurun_out
R=${1:-r02a}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logpost_kernel -s 2 -c 1 -f -o gpurun_out/logpost_$R \
   python scripts/gpu_time_engine.py 189440 > gpurun_out/ncu_logpost_$R.log 2>&1; echo "ncu logpost rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbfgs_kernel -c 1 -f -o gpurun_out/lbfgs_$R \
   python scripts/gpu_time_map.py 2368 300 > gpurun_out/ncu_lbfgs_$R.log 2>&1; echo "ncu lbfgs rc=$?"
ls -la gpurun_out | tail
