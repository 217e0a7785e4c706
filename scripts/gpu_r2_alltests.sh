#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q 2>&1 | tail -25 | cut -c1-300 > gpurun_out/r2_alltests.log
cat gpurun_out/r2_alltests.log
