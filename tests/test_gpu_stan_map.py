"""GPU: the CUDA path against REAL Stan output -- the MAP fits the reference's authors saved for their paper (fixture
tests/golden/stan_map.npz, see test_oracle_stan_map.py).  At the parameter values ``StanModel.optimizing`` returned, with
kernel / penalty matrices built by the CUDA kernels for the stored grids:

  * ``bdrt_constrain`` reproduces Stan's transformed parameters (sigma_tot -- which contains Z_hat --, R_inf, xp) to 1e-9
    for all 25 fits: 'Series', 'Series_pos', 'Series-Parallel_pos' and 'Series-2Parallel_pos' programs;
  * the fused log density / gradient equals the oracle's there (1e-11 / 1e-9), so the stationarity statements of
    test_oracle_stan_map.py carry over;
  * where Stan converged tightly, ``bdrt_map_newton`` started from Stan's end point stays there: x within 1e-2 of its
    peak, log density gain below 0.1."""
import numpy as np
import pytest
import torch

from oracle import model as omod, model_sp as osp
from test_oracle_stan_map import CPU_ONLY, NAMES, TIGHT, build

pytestmark = pytest.mark.gpu


def gpu_problem_any(name, d, meta):
    """capi.SeriesProblem for one oracle data dict of any family, every matrix built by the CUDA path."""
    from bayes_drt_b200 import capi
    f = torch.tensor(d['freq'])
    Zt = torch.tensor(d['Z'][None, :])
    if 'As' not in d:
        tau = torch.tensor(d['tau'])
        A = torch.cat(capi.build_A(f, tau, d['epsilon']))
        c = omod.MODE_CONSTANTS['optimize']
        bf = torch.tensor(1 / (2 * np.pi * d['tau']))
        L = torch.stack([c[f'l{o}'] * capi.build_L(bf, tau, d['epsilon'], o) for o in (0, 1, 2)])
        return capi.SeriesProblem(A, Zt, f, L, nonneg=d['pos'], sigma_min=d['sigma_min'], ups_alpha=d['ups_alpha'],
                                  ups_beta=d['ups_beta'], induc_scale=d['induc_scale'])
    c = osp.MODE_CONSTANTS_SP['optimize']

    def mats(tau, eps, key, info=None):
        t = torch.tensor(tau)
        kw = {} if info is None else dict(kernel='DDT', dist_type='parallel', symmetry=info['symmetry'], bc=info['bc'])
        A = torch.cat(capi.build_A(f, t, eps, **kw))
        L = torch.stack([c[key][o] * capi.build_L(torch.tensor(1 / (2 * np.pi * tau)), t, eps, o) for o in range(3)])
        return A, L
    As, Ls = mats(d['tau_s'], d['eps_s'], 'ls')
    pars = meta['pars']
    Ap, Lp = mats(d['tau_p'], d['eps_p'], 'lp', pars[0])
    kw = dict(Ap=Ap, Lp=Lp, x_sum_invscale=d['x_sum_invscale'], xp_scale=d['xp_scale'])
    if 'Ap2' in d:
        Ap2, Lp2 = mats(d['tau_p2'], d['eps_p2'], 'lp', pars[1])
        kw.update(Ap2=Ap2, Lp2=Lp2, xp2_scale=d['xp2_scale'])
    return capi.SeriesProblem(As, Zt, f, Ls, nonneg=True, sigma_min=d['sigma_min'], ups_alpha=d['ups_alpha'],
                              ups_beta=d['ups_beta'], induc_scale=d['induc_scale'], **kw)


@pytest.mark.parametrize('name', [n for n in NAMES if n not in CPU_ONLY])
def test_cuda_reproduces_stans_transformed_parameters(name):
    d, u, S, mod, meta = build(name)
    prob = gpu_problem_any(name, d, meta)
    assert prob.D == len(u)
    out = prob.split_outputs(prob.constrain(torch.tensor(u[None, :])))
    sig = out['sigma_tot'][0].cpu().numpy()
    assert np.max(np.abs(sig - S['sigma_tot']) / S['sigma_tot']) <= 1e-9, name
    assert abs(out['Rinf'][0].item() - S['Rinf']) <= 1e-12 * S['Rinf']
    if mod is omod:
        assert np.allclose(out['x'][0].cpu().numpy(), S['x'], rtol=1e-12, atol=0)
    else:
        assert np.allclose(out['xs'][0].cpu().numpy(), S['xs'], rtol=1e-12, atol=0)
        for q in (('xp',) if 'Ap2' not in d else ('xp1', 'xp2')):
            if q in S:
                assert np.allclose(out[q][0].cpu().numpy(), S[q], rtol=1e-12, atol=0), q
    lp, grad = prob.logpost_grad(torch.tensor(u[None, :]))
    lo, go = mod.logpost(u, d)
    assert abs(lp[0].item() - lo) <= 1e-11 * abs(lo), (lp[0].item(), lo)
    gerr = np.max(np.abs(grad[0].cpu().numpy() - go))
    # (Stan's end points are near-stationary: the net gradient is ~1 while its terms are ~|lp|, so the rounding error is
    # judged on the scale of the terms)
    assert gerr <= 1e-9 * np.max(np.abs(go)) + 2e-10 * abs(lo), (gerr, np.max(np.abs(go)), int(np.argmax(np.abs(grad[0].cpu().numpy() - go))))


@pytest.mark.parametrize('name', TIGHT)
def test_cuda_newton_from_stans_optimum_stays_there(name):
    d, u, S, mod, meta = build(name)
    prob = gpu_problem_any(name, d, meta)
    lp0, _ = prob.logpost_grad(torch.tensor(u[None, :]))
    r = prob.map_newton(torch.tensor(u[None, :]))
    gain = r['lp'][0].item() - lp0[0].item()
    assert 0.0 <= gain <= 0.1, gain
    x = prob.split_outputs(prob.constrain(r['u']))['x'][0].cpu().numpy()
    assert np.max(np.abs(x - S['x'])) <= 1e-2 * np.max(np.abs(S['x']))
