#!/bin/bash
# A/B of the synchronised warp mode (BDRT_WSYNC=1: one CTA barrier at the entry of every evaluation) on one box
mkdir -p gpurun_out
{
for rep in 1 2; do
for ws in 0 1; do
  export BDRT_WSYNC=$ws
  echo "== nuts wsync=$ws"; timeout 300 python scripts/gpu_time_nuts.py 1184 2 200 200 2>&1 | grep "^B="
done
done
unset BDRT_WSYNC
echo "== map coop (default)"; timeout 300 python scripts/gpu_time_map.py 12500 50000 2>&1 | grep "^B=" | tail -1
export BDRT_WARP=1
for ws in 0 1; do
  export BDRT_WSYNC=$ws
  echo "== map warp wsync=$ws"; timeout 300 python scripts/gpu_time_map.py 12500 50000 2>&1 | grep "^B=" | tail -1
  echo "== map uniform warp wsync=$ws"; timeout 300 python scripts/gpu_time_map.py 4736 2000 2>&1 | grep "^B=" | tail -1
done
unset BDRT_WARP
export BDRT_WSYNC=1
echo "== tests with BDRT_WSYNC=1"
timeout 900 python -m pytest tests/test_gpu_nuts.py tests/test_gpu_map.py tests/test_gpu_per_spectrum.py -m gpu -q 2>&1 | tail -4
} > gpurun_out/r2_wsync.log 2>&1
cat gpurun_out/r2_wsync.log
