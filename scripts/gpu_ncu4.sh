#!/bin/bash
R=${1:-r01e}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts_kernel -c 1 -f -o gpurun_out/nuts_$R \
   python scripts/gpu_time_nuts.py 1184 2 24 8 > gpurun_out/ncu_nuts_$R.log 2>&1; echo "ncu nuts rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbfgs_kernel -c 1 -f -o gpurun_out/lbfgs_$R \
   python scripts/gpu_time_map.py 2368 300 > gpurun_out/ncu_lbfgs_$R.log 2>&1; echo "ncu lbfgs rc=$?"
