#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on shared memory) over a slice of the GPU tests and over a small run of every
# solver on every operand layout of the engine.  Output: gpurun_out/sanitize_*.log and a summary on stdout.
# `bash scripts/gpu_sanitize.sh race` runs the racecheck part only.
mkdir -p gpurun_out
ONLY=$1
if [ "$ONLY" != race ]; then
T="tests/test_gpu_logpost.py::test_logpost_S_shape tests/test_gpu_logpost.py::test_logpost_non_toeplitz_grid tests/test_gpu_series_parallel.py::test_sp_logpost_matches_oracle tests/test_gpu_series_parallel.py::test_sp_dense_too_large_goes_global tests/test_gpu_summaries.py tests/test_gpu_matrices.py tests/test_gpu_ridge.py::test_qp_bound_random_problems tests/test_gpu_cvxopt_ridge.py tests/test_gpu_stan_map.py::test_cuda_newton_from_stans_optimum_stays_there"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -x -q -m gpu > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck tests rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_memcheck.log | head -20
fi
cat > /tmp/san_small.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np, torch
from helpers import gpu_problem, load_spectrum, oracle_batch
from bayes_drt_b200 import Inverter, capi
freq, Z = load_spectrum('ZARC_uniform_0.25')
for kw in (dict(), dict(nonneg=True, outliers=True)):
    ds = oracle_batch(freq, [Z, Z], mode='optimize', **kw)
    prob = gpu_problem(ds)
    u0 = torch.tensor(np.random.RandomState(0).uniform(-2, 2, (2, prob.D)))
    r = prob.map_lbfgs(u0, max_iter=8)
    p = prob.map_newton(r['u'], max_iter=2)
    ds2 = oracle_batch(freq, [Z, Z], mode='sample', **kw)
    prob2 = gpu_problem(ds2)
    g = torch.Generator().manual_seed(0)
    n = prob2.nuts(torch.rand(2, 2, prob2.D, generator=g, dtype=torch.float64) * 4 - 2, chains=2, warmup=3, samples=2, max_treedepth=4)
    torch.cuda.synchronize()
    print('ok', kw, r['lp'].tolist(), torch.isfinite(n['draws']).all().item())
# per-spectrum grids through the facade (per-slot tables in warp mode; dense per-spectrum grids when forced) + ridge
fr = torch.tensor(np.stack([freq, freq * 1.1]))
Zb = torch.tensor(np.stack([Z, Z]))
inv = Inverter()
inv.fit(fr, Zb, mode='optimize', max_iter=6, check_outliers=False)
inv.ridge_fit(fr, Zb, max_iter=3)
torch.cuda.synchronize()
print('ok per-spectrum grids', inv.R_inf.tolist())
PY
if [ "$ONLY" != race ]; then
for mode in default warp dense gdense; do
  unset BDRT_WARP BDRT_FORCE_DENSE BDRT_FORCE_GDENSE
  [ $mode = warp ] && export BDRT_WARP=1
  [ $mode = dense ] && export BDRT_FORCE_DENSE=1
  [ $mode = gdense ] && export BDRT_FORCE_DENSE=1 BDRT_FORCE_GDENSE=1
  timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_small.py > gpurun_out/sanitize_solvers_$mode.log 2>&1; echo "memcheck solvers $mode rc=$?"; grep -E "ERROR SUMMARY|^ok|Invalid|out of bounds" gpurun_out/sanitize_solvers_$mode.log | head -6
done
fi
for mode in default warp; do
  unset BDRT_WARP BDRT_FORCE_DENSE BDRT_FORCE_GDENSE
  [ $mode = warp ] && export BDRT_WARP=1
  timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san_small.py > gpurun_out/sanitize_race_$mode.log 2>&1; echo "racecheck $mode rc=$?"; grep -E "RACECHECK SUMMARY|^ok|hazard" gpurun_out/sanitize_race_$mode.log | head -6
done
