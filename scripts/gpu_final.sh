#!/bin/bash
mkdir -p gpurun_out
R=${1:-r01h}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$R.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench rc=$?"; cat gpurun_out/bench_$R.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_$R.json
timeout 600 python scripts/gpu_time_config5.py 296 8 > gpurun_out/config5_$R.log 2>&1; cat gpurun_out/config5_$R.log
