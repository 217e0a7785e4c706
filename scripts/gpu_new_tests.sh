#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest ${@:-tests/test_gpu_ridge.py} -x -q -m gpu 2>&1 | tail -30 > gpurun_out/new_tests.log
cat gpurun_out/new_tests.log
