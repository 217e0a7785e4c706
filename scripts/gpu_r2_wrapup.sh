#!/bin/bash
# final state of the round: whole GPU suite, smoke, both bench arms (same commands as the driver's)
mkdir -p gpurun_out
R=${1:-r02h}
timeout 3000 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300 > gpurun_out/${R}_pytest_gpu.txt; cat gpurun_out/${R}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_$R.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref_$R.json
