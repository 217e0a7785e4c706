#!/bin/bash
mkdir -p gpurun_out
{
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_map_benchmark.py -x -q -s 2>&1 | grep "MAP parity\|passed\|failed\|Error" | cut -c1-1500
timeout 600 python -m pytest tests/test_gpu_inverter.py -x -q 2>&1 | tail -4
} > gpurun_out/r2_dbg.log 2>&1
cat gpurun_out/r2_dbg.log
