"""GPU: MAP parity on the benchmark shape (BASELINE.json config 4: Nf = 70, K = 100, model 'Series', shared grid) through
the drop-in call ``Inverter.fit(freq, Z, mode='optimize')`` on 256 synthetic spectra.

North star: MAP DRT coefficients within 1e-5 relative of the optimiser.  Stan's L-BFGS stops on its relative tests
1e-3 .. 4e-2 short of the optimum (SURVEY section 7 hard part 1; measured below), so the 1e-5 statement is made about
``fit(..., polish=True)`` against the oracle's own Newton-converged optimum started from the same L-BFGS end point
(oracle/newton.py: central-difference Hessian, numpy Cholesky -- nothing shared with csrc/newton.cu), and the distance
of the unpolished default from that optimum is measured and reported next to it."""
import json
import multiprocessing as mp
import os

import numpy as np
import pytest
import torch

from helpers import oracle_batch
from oracle import model as omod, newton as onew

pytestmark = pytest.mark.gpu
NSPEC = 256
_DS = None


def _func(d):
    def f(u):
        lp, g = omod.logpost(u, d)
        if not np.isfinite(lp) or not np.all(np.isfinite(g)):
            return None
        return -lp, -g
    return f


def _oracle_polish(args):
    b, u = args
    with np.errstate(all='ignore'):
        o = onew.polish(_func(_DS[b]), u, max_iter=120)
    c = omod.constrain(o['x'], _DS[b])
    return dict(x=c['x'], Rinf=c['Rinf'], sigma_res=c['sigma_res'], f=o['f'], gnorm=o['gnorm'], failed=o['failed'])


def test_map_1e5_on_benchmark_batch():
    global _DS
    from bayes_drt_b200 import Inverter, synth
    freq, Z, _ = synth.make_spectra(NSPEC, seed=20240601)
    _, bf = synth.bench_grid()
    # unpolished default (what bench.py times) and the polished fit, same hash-keyed random starts
    inv = Inverter(basis_freq=bf.numpy())
    inv.fit(freq, Z, mode='optimize', check_outliers=False)
    u_lbfgs = inv._opt_result['u'].cpu().numpy()
    s = inv._Z_scale.cpu().numpy()
    x_lbfgs = inv.distribution_fits['DRT']['coef'].cpu().numpy() / s[:, None]
    status = inv._opt_result['status'].cpu().numpy()
    assert np.isin(status, (10, 20, 21, 30, 31, 40)).all(), np.unique(status)
    pol = Inverter(basis_freq=bf.numpy())
    pol.fit(freq, Z, mode='optimize', polish=True, check_outliers=False)
    x_pol = pol.distribution_fits['DRT']['coef'].cpu().numpy() / s[:, None]
    Rinf_pol = pol.R_inf.cpu().numpy() / s
    sres_pol = pol.error_fit['sigma_res'].cpu().numpy() / s
    gnorm = pol._opt_result['gnorm'].cpu().numpy()
    assert np.array_equal(pol._opt_result['iters'].cpu().numpy(), inv._opt_result['iters'].cpu().numpy())

    # oracle: Newton from the same L-BFGS end points, all host cores (fork: the workers only run numpy)
    _DS = oracle_batch(freq.numpy(), list(Z.numpy()), basis_freq=bf.numpy(), mode='optimize')
    assert np.allclose([d['Z_scale'] for d in _DS], s, rtol=1e-12)
    with mp.get_context('fork').Pool(min(os.cpu_count() or 1, 32)) as pool:
        ora = pool.map(_oracle_polish, [(b, u_lbfgs[b]) for b in range(NSPEC)], chunksize=4)
    # converged = max|grad| within a factor 2 of the 1e-9 both Newton iterations aim for (measured on this batch: with
    # condition numbers of 3e7 and more, a spectrum stuck at |grad| = 5e-9 is still 4e-4 of the peak from the optimum
    # along its flattest direction, while every spectrum below 2e-9 agrees to 1e-10) -- the oracle's `failed` flag only
    # says that its last damped step could not improve on a point already at the rounding floor
    ogn = np.array([o['gnorm'] for o in ora])
    ok = (ogn < 2e-9) & (gnorm < 2e-9)
    err_pol, err_lbfgs, err_R, err_s, err_lp = (np.zeros(NSPEC) for _ in range(5))
    for b in range(NSPEC):
        xo = ora[b]['x']
        sc = np.max(np.abs(xo))
        err_pol[b] = np.max(np.abs(x_pol[b] - xo)) / sc
        err_lbfgs[b] = np.max(np.abs(x_lbfgs[b] - xo)) / sc
        # scalars: relative, with an absolute floor on the scale of sigma_min = 0.002 for the lower=0 parameters that go to
        # their bound (sigma_res for clean RC spectra; R_inf in a few poor optima)
        err_R[b] = abs(Rinf_pol[b] - ora[b]['Rinf']) / (ora[b]['Rinf'] + 0.2)
        err_s[b] = abs(sres_pol[b] - ora[b]['sigma_res']) / (ora[b]['sigma_res'] + 0.2)
        err_lp[b] = abs(pol._opt_result['lp'][b].item() + ora[b]['f']) / abs(ora[b]['f'])
    q = lambda a: {p: float(np.percentile(a[ok], p)) for p in (5, 50, 95, 100)}  # noqa: E731
    rep = dict(n=NSPEC, converged_both=int(ok.sum()), converged_cuda=int((gnorm < 2e-9).sum()),
               converged_oracle=int((ogn < 2e-9).sum()), polished_rel_err_inf=q(err_pol), lbfgs_rel_err_inf=q(err_lbfgs),
               termination={int(k): int((status == k).sum()) for k in np.unique(status)},
               worst=[dict(b=int(b), err=float(err_pol[b]), gnorm_cuda=float(gnorm[b]), gnorm_oracle=float(ogn[b]))
                      for b in np.argsort(-np.where(ok, err_pol, 0))[:5]])
    print('MAP parity on the benchmark shape:', json.dumps(rep))
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/map_parity_benchmark.json', 'w') as fh:
        json.dump(rep, fh, indent=1)
    # both Newton iterations converge for (nearly) every spectrum, poor local optima of the random starts included
    assert ok.mean() >= 0.95, rep
    assert err_pol[ok].max() <= 1e-5, rep           # north star: MAP coefficients within 1e-5 (of the largest one)
    assert err_R[ok].max() <= 1e-5 and err_s[ok].max() <= 1e-5, (err_R[ok].max(), err_s[ok].max())
    assert err_lp[ok].max() <= 1e-9
    # the unpolished default is Stan's own termination: percent-level away from the optimum, never 1e-5
    assert 1e-5 < np.median(err_lbfgs[ok]) < 0.2
