#!/bin/bash
mkdir -p gpurun_out
R=${1:-r01g}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
   python bench.py --steps 2 --warmup 1 --cpu-sample 8 > gpurun_out/bench_under_ncu_$R.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbfgs_kernel -c 1 -f -o gpurun_out/lbfgs_$R \
   python scripts/gpu_time_map.py 2368 300 > gpurun_out/ncu_lbfgs_$R.log 2>&1; echo "ncu lbfgs rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts_kernel -c 1 -f -o gpurun_out/nuts_$R \
   python scripts/gpu_time_nuts.py 1184 2 24 8 > gpurun_out/ncu_nuts_$R.log 2>&1; echo "ncu nuts rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:summarize_kernel -c 1 -f -o gpurun_out/summarize_$R \
   python scripts/gpu_time_summarize.py > gpurun_out/ncu_summarize_$R.log 2>&1; echo "ncu summarize rc=$?"
timeout 600 python scripts/gpu_time_config5.py 296 4 > gpurun_out/config5_$R.log 2>&1; cat gpurun_out/config5_$R.log
du -sh gpurun_out
