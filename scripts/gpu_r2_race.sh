#!/bin/bash
# after a change of the engine's per-slot phases: racecheck, the parity tests that are sensitive to it, and the two timings
bash scripts/gpu_sanitize.sh race
timeout 900 python -m pytest tests/test_gpu_logpost.py tests/test_gpu_map.py tests/test_gpu_nuts.py::test_nuts_deterministic_and_shard_independent tests/test_gpu_series_parallel.py -m gpu -q 2>&1 | tail -3 | cut -c1-300
timeout 300 python scripts/gpu_time_map.py 12500 50000 2>&1 | grep "^B=" | tail -1
timeout 300 python scripts/gpu_time_nuts.py 1184 2 200 200 2>&1 | grep "^B="
