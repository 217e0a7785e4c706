#!/bin/bash
# End-of-round evidence: smoke, bench (ours + reference arm), ncu launch list of the bench command, one --set full
# capture of each solver kernel.  Outputs under gpurun_out/ (summarised into profiles/ by scripts/ncu_summary.py).
mkdir -p gpurun_out
R=${1:-r01}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$R.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench rc=$?"; cat gpurun_out/bench_$R.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_$R.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
   python bench.py --steps 2 --warmup 1 --cpu-sample 8 > gpurun_out/bench_under_ncu_$R.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbfgs_kernel -c 1 -f -o gpurun_out/lbfgs_$R \
   python scripts/gpu_time_map.py 2368 300 > gpurun_out/ncu_lbfgs_$R.log 2>&1; echo "ncu lbfgs rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts_kernel -c 1 -f -o gpurun_out/nuts_$R \
   python scripts/gpu_time_nuts.py 1184 2 24 8 > gpurun_out/ncu_nuts_$R.log 2>&1; echo "ncu nuts rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logpost_kernel -s 2 -c 1 -f -o gpurun_out/logpost_$R \
   python scripts/gpu_time_engine.py 189440 > gpurun_out/ncu_logpost_$R.log 2>&1; echo "ncu logpost rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:summarize_kernel -c 1 -f -o gpurun_out/summarize_$R \
   python scripts/gpu_time_summarize.py > gpurun_out/ncu_summarize_$R.log 2>&1; echo "ncu summarize rc=$?"
timeout 300 python scripts/gpu_time_summarize.py > gpurun_out/summarize_$R.log 2>&1; cat gpurun_out/summarize_$R.log
ls -la gpurun_out | tail -30
