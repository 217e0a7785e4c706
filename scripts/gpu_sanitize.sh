#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on shared memory) over a small slice of the GPU tests
mkdir -p gpurun_out
T="tests/test_gpu_logpost.py::test_logpost_S_shape tests/test_gpu_logpost.py::test_logpost_non_toeplitz_grid tests/test_gpu_series_parallel.py::test_sp_logpost_matches_oracle tests/test_gpu_summaries.py::test_summarize_matches_numpy tests/test_gpu_matrices.py tests/test_gpu_ridge.py::test_qp_bound_random_problems"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -x -q > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_memcheck.log | head -20
cat > /tmp/san_small.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np, torch
from helpers import gpu_problem, load_spectrum, oracle_batch
freq, Z = load_spectrum('ZARC_uniform_0.25')
for kw in (dict(), dict(nonneg=True, outliers=True)):
    ds = oracle_batch(freq, [Z, Z], mode='optimize', **kw)
    prob = gpu_problem(ds)
    u0 = torch.tensor(np.random.RandomState(0).uniform(-2, 2, (2, prob.D)))
    r = prob.map_lbfgs(u0, max_iter=8)
    p = prob.map_newton(r['u'], max_iter=1)
    ds2 = oracle_batch(freq, [Z, Z], mode='sample', **kw)
    prob2 = gpu_problem(ds2)
    g = torch.Generator().manual_seed(0)
    n = prob2.nuts(torch.rand(2, 2, prob2.D, generator=g, dtype=torch.float64) * 4 - 2, chains=2, warmup=3, samples=2, max_treedepth=4)
    torch.cuda.synchronize()
    print('ok', kw, r['lp'].tolist(), torch.isfinite(n['draws']).all().item())
PY
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_small.py > gpurun_out/sanitize_solvers.log 2>&1; echo "memcheck solvers rc=$?"; grep -E "ERROR SUMMARY|^ok|Invalid|out of bounds" gpurun_out/sanitize_solvers.log | head
BDRT_FORCE_DENSE=1 timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_small.py > gpurun_out/sanitize_solvers_dense.log 2>&1; echo "memcheck solvers dense rc=$?"; grep -E "ERROR SUMMARY|^ok|Invalid|out of bounds" gpurun_out/sanitize_solvers_dense.log | head
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san_small.py > gpurun_out/sanitize_race.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|^ok|hazard" gpurun_out/sanitize_race.log | head
