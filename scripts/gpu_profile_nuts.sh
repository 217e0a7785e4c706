#!/bin/bash
# source-level sampling of nuts_kernel in its steady state (long trees): few replays, so a long kernel is affordable
mkdir -p gpurun_out
R=${1:-r01i}
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SpeedOfLight --clock-control none --import-source on \
   -k regex:nuts_kernel -c 1 -f -o gpurun_out/nuts_$R python scripts/gpu_time_nuts.py 1184 2 120 10 > gpurun_out/ncu_nuts_$R.log 2>&1
echo "ncu nuts rc=$?"; tail -5 gpurun_out/ncu_nuts_$R.log; du -sh gpurun_out
