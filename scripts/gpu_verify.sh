#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_final.log
timeout 330 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_final.log; cat gpurun_out/pytest_final.log
