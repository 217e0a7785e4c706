#!/bin/bash
mkdir -p gpurun_out
{ BDRT_LIB=$PWD/scratch_libs/libbdrt_clk.so timeout 300 python scripts/gpu_phase_clocks.py 2>&1 | tail -8
  echo "== uniform work"; timeout 300 python scripts/gpu_time_map.py 4736 2000 2>&1 | grep "^B=\|status"; } > gpurun_out/r2_clk.log 2>&1
cat gpurun_out/r2_clk.log
