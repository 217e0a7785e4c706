#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + one --set full capture of each solver kernel
mkdir -p gpurun_out
R=${1:-r01}
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench rc=$?"; cat gpurun_out/bench_$R.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_$R.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
   python bench.py --steps 2 --warmup 1 --no-hmc --cpu-sample 8 > gpurun_out/bench_under_ncu_$R.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbfgs_kernel -c 1 -f -o gpurun_out/lbfgs_$R \
   python scripts/gpu_time_map.py 1184 150 > gpurun_out/ncu_lbfgs_$R.log 2>&1; echo "ncu lbfgs rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logpost_kernel -s 2 -c 1 -f -o gpurun_out/logpost_$R \
   python scripts/gpu_time_engine.py 94720 > gpurun_out/ncu_logpost_$R.log 2>&1; echo "ncu logpost rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nuts_kernel -c 1 -f -o gpurun_out/nuts_$R \
   python scripts/gpu_time_nuts.py 592 2 6 4 > gpurun_out/ncu_nuts_$R.log 2>&1; echo "ncu nuts rc=$?"
ls -la gpurun_out
