#!/bin/bash
mkdir -p gpurun_out
R=${1:-r02c}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$R.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench rc=$?"; cat gpurun_out/bench_$R.json; tail -5 gpurun_out/bench_$R.err
