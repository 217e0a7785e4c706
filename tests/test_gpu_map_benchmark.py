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
    """Oracle optimum from the L-BFGS end point u, and Newton's own error estimate of both polished points: the largest
    change of a compared quantity by one more (undamped) Newton step -- |H^-1 g|_inf over x relative to max|x|, and the
    change of R_inf and sigma_res on the scale they are compared on below -- with the oracle's gradient at either point
    and the oracle's central-difference Hessian at its optimum.  (Coordinates sent to the floor of
    a lower=0 constraint are held; a point with a coordinate at the floor whose d lp / d theta is positive is not a
    constrained optimum at all and gets an infinite estimate -- the sign of that derivative survives the factor theta.)  The rounding floor of the gradient is
    1e-10 .. 1e-8 depending on the spectrum, and what it means for x is decided by the flattest direction of H --
    max|g| alone does not say."""
    b, u, u_cuda = args
    f = _func(_DS[b])
    with np.errstate(all='ignore'):
        o = onew.polish(f, u, max_iter=120)
        H = onew.fd_hessian(lambda z: f(z)[1], o['x'])
        held = (o['x'] < -30) | (u_cuda < -30)
        H[held, :] = 0
        H[:, held] = 0
        H[held, held] = 1
        ps = omod.param_slices(_DS[b])
        sl = ps['x']
        est = []
        for pt in (o['x'], u_cuda):
            r = f(pt)
            if r is None or np.any(r[1][pt < -30] < 0):  # (r[1] is the gradient of f = -lp)
                est.append(np.inf)
                continue
            g = np.where(held, 0.0, r[1])
            try:
                st = np.linalg.solve(H, g)
            except np.linalg.LinAlgError:
                est.append(np.inf)
                continue
            e = float(np.max(np.abs(st[sl])) / np.max(np.abs(o['x'][sl])))
            cpt = omod.constrain(pt, _DS[b])
            for raw, name in (('Rinf_raw', 'Rinf'), ('sigma_res_raw', 'sigma_res')):  # theta ~ exp(u): d theta = theta du
                e = max(e, float(abs(st[ps[raw]][0]) * cpt[name] / (cpt[name] + 0.2)))
            est.append(e)
    c = omod.constrain(o['x'], _DS[b])
    return dict(x=c['x'], Rinf=c['Rinf'], sigma_res=c['sigma_res'], f=o['f'], gnorm=o['gnorm'], failed=o['failed'],
                est_oracle=est[0], est_cuda=est[1])


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
    u_pol = pol._opt_result['u'].cpu().numpy()
    assert np.array_equal(pol._opt_result['iters'].cpu().numpy(), inv._opt_result['iters'].cpu().numpy())

    # oracle: Newton from the same L-BFGS end points, all host cores (fork: the workers only run numpy)
    _DS = oracle_batch(freq.numpy(), list(Z.numpy()), basis_freq=bf.numpy(), mode='optimize')
    assert np.allclose([d['Z_scale'] for d in _DS], s, rtol=1e-12)
    with mp.get_context('fork').Pool(min(os.cpu_count() or 1, 32)) as pool:
        ora = pool.map(_oracle_polish, [(b, u_lbfgs[b], u_pol[b]) for b in range(NSPEC)], chunksize=4)
    # converged = Newton's error estimate of BOTH end points below 1e-6 (a tenth of the tolerance asserted below).  A
    # threshold on max|grad| does not work: both iterations aim for 1e-9, the gradient's rounding floor is above that for
    # some 15 % of the spectra (they stop at 2e-9 .. 1e-8 and still agree to 1e-10), while one spectrum of this batch
    # with a condition number far above the typical 3e7 is 4e-4 of the peak from its optimum at max|grad| = 5e-9.  The
    # oracle's `failed` flag only says that its last damped step could not improve on a point at the rounding floor.
    ogn = np.array([o['gnorm'] for o in ora])
    est_o = np.array([o['est_oracle'] for o in ora])
    est_c = np.array([o['est_cuda'] for o in ora])
    ok = (est_o < 1e-6) & (est_c < 1e-6)
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
    rep = dict(n=NSPEC, converged_both=int(ok.sum()), converged_cuda=int((est_c < 1e-6).sum()),
               converged_oracle=int((est_o < 1e-6).sum()), gnorm_cuda_below_2e9=int((gnorm < 2e-9).sum()),
               gnorm_oracle_below_2e9=int((ogn < 2e-9).sum()),
               newton_iterations_cuda=({p: float(np.percentile(pol._opt_result['newton_iters'].cpu().numpy(), p))
                                        for p in (50, 95, 100)} if 'newton_iters' in pol._opt_result else None),
               polished_rel_err_inf=q(err_pol), lbfgs_rel_err_inf=q(err_lbfgs),
               termination={int(k): int((status == k).sum()) for k in np.unique(status)},
               worst=[dict(b=int(b), err=float(err_pol[b]), gnorm_cuda=float(gnorm[b]), gnorm_oracle=float(ogn[b]),
                           est_cuda=float(est_c[b]), est_oracle=float(est_o[b]))
                      for b in np.argsort(-np.where(ok, err_pol, 0))[:5]],
               not_determined=[dict(b=int(b), err=float(err_pol[b]), gnorm_cuda=float(gnorm[b]),
                                    gnorm_oracle=float(ogn[b]), est_cuda=float(est_c[b]), est_oracle=float(est_o[b]))
                               for b in np.where(~ok)[0][:10]])
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
