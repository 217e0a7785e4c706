#!/bin/bash
# End-of-round evidence: smoke, bench (ours + reference arm), ncu launch list + DRAM traffic of the bench command, one
# --set full capture of every kernel of the path.  The .ncu-rep files are summarised ON THE BOX (profiles/<tag>_*.txt,
# profiles/summary.json) and only the summaries are copied to gpurun_out/ (gpurun returns at most 64 MiB); of the reports
# only the L-BFGS one is kept.  Usage: bash scripts/gpu_r2_final.sh <tag> [skipbench]
mkdir -p gpurun_out
R=${1:-r02}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$R.log
# DRAM traffic of lbfgs_kernel on the bench's own launch shape first, so that the bench line below carries it
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches_$R.csv python bench.py --quick --steps 1 --warmup 1 --cpu-sample 8 > gpurun_out/bench_under_ncu_$R.log 2>&1; echo "ncu list rc=$?"
python scripts/ncu_traffic.py $R gpurun_out/launches_$R.csv 12500 50000
python scripts/launch_summary.py gpurun_out/launches_$R.csv gpurun_out/${R}_launches_summary.txt > /dev/null
gzip -f gpurun_out/launches_$R.csv
if [ "$2" != "skipbench" ]; then
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$R.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_$R.json
fi
cap() {  # name, ncu kernel options, command...
  local name=$1 opts=$2; shift 2
  timeout 900 ncu --set full --clock-control none --import-source on $opts -c 1 -f -o /tmp/${name}_$R "$@" > gpurun_out/ncu_${name}_$R.log 2>&1
  echo "ncu $name rc=$?"
  python scripts/ncu_summary.py $R ${name}_kernel=/tmp/${name}_$R.ncu-rep
  if [ "$name" = lbfgs ] || [ "$name" = nuts ]; then python scripts/ncu_phases.py /tmp/${name}_$R.ncu-rep > profiles/${R}_${name}_phases.txt 2>&1; fi
}
cap lbfgs "-k regex:lbfgs_kernel" python scripts/gpu_time_map.py 2368 300
cap nuts "-k regex:nuts_kernel" python scripts/gpu_time_nuts.py 1184 2 24 8
cap logpost "-k regex:logpost_kernel -s 2" python scripts/gpu_time_engine.py 189440
cap ridge "-k regex:ridge_kernel -s 2" python scripts/gpu_time_ridge.py 2048
cap build_A "-k regex:build_A_kernel -s 1" python scripts/gpu_time_misc.py A
cap newton "-k regex:newton_kernel" python scripts/gpu_time_misc.py newton
cp profiles/${R}_* profiles/summary.json gpurun_out/
cp /tmp/lbfgs_$R.ncu-rep gpurun_out/ 2>/dev/null
timeout 300 python scripts/gpu_time_misc.py newton 2>&1 | tail -1 | tee gpurun_out/newton_time_$R.log
timeout 300 python scripts/gpu_time_ridge.py 8192 2>&1 | tail -4 | tee gpurun_out/ridge_time_$R.log
du -sh gpurun_out
