"""GPU: statistical parity of the batched NUTS sampler for BASELINE.json configs 3 and 5 against long runs of the oracle's
restatement of Stan's sampler (tests/golden/nuts_config_{outliers,sp}.npz from scripts/make_golden_nuts_configs.py:
16 chains x (300 + 2000), split R-hat <= 1.007, bulk ESS >= 3500 for every compared quantity):

  config 3  Stan program Series_outliers (inversion.py:1218-1221 with outliers=True; Series_outliers_modelcode.txt:45-72),
            ZARC spectrum with three gross outliers, D = 373: x, R_inf, inductance, error-model scalars and sigma_out[Nf]
  config 5  Stan program Series-Parallel_pos (Series-Parallel_modelcode.txt:52-107) on the reference's
            data/simulated/Z_DRT-2-TpDDT_uniform_0.25.csv with the paper's two distributions, D = 336: xs, xp, R_inf, ...

North star: posterior means and the end points of the 95 % interval within 3 Monte-Carlo standard errors (MCSE of both
sides combined).  Random streams cannot match (Philox vs numpy), so parity is statistical."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLD, gpu_problem, gpu_problem_sp, oracle_batch

pytestmark = pytest.mark.gpu


def _compare(cons, gold, scale_idx):
    """cons [chains, samples, P] constrained draws of the CUDA sampler; gold: the oracle's summary file."""
    from oracle.nuts import mcse_mean, mcse_quantile
    P = cons.shape[-1]
    flat = cons.reshape(-1, P)
    mean = flat.mean(0)
    q025, q975 = np.percentile(flat, 2.5, axis=0), np.percentile(flat, 97.5, axis=0)
    se_mean = np.hypot([mcse_mean(cons[:, :, i]) for i in range(P)], gold['mcse_mean'])
    se_lo = np.hypot([mcse_quantile(cons[:, :, i], 0.025) for i in range(P)], gold['mcse_q025'])
    se_hi = np.hypot([mcse_quantile(cons[:, :, i], 0.975) for i in range(P)], gold['mcse_q975'])
    # quantities that are ~0 with a tiny spread (coefficients far in the tails of a distribution) are also judged on the
    # scale of their group's largest posterior mean: a difference below 2e-3 of it is not a violation
    floor = np.zeros(P)
    for idx in scale_idx:
        floor[idx] = 2e-3 * np.abs(gold['mean'][idx]).max()
    report = {}
    for name, a, b, se in (('mean', mean, gold['mean'], se_mean), ('q025', q025, gold['q025'], se_lo),
                           ('q975', q975, gold['q975'], se_hi)):
        z = np.abs(a - b) / se
        viol = (z > 3.0) & (np.abs(a - b) > floor)
        report[name] = (float(np.median(z)), float(viol.mean()), float(z[viol].max()) if viol.any() else 0.0)
        # 3 x ~200 comparisons with estimated MCSEs: a few excursions beyond 3 are expected by chance (0.27 % each for an
        # exact normal); the systematic-error alarm is the fraction beyond 3 and anything beyond 6
        assert viol.mean() <= 0.02, (name, np.where(viol)[0], z[viol])
        assert not viol.any() or z[viol].max() < 6.0, (name, np.where(viol)[0], z[viol])
        assert np.median(z) < 1.5, (name, np.median(z))  # unbiased: |z| of a standard normal has median 0.67
    return report


def _nuts(prob, chains, warmup, samples, seed):
    g = torch.Generator().manual_seed(seed)
    u0 = torch.rand(prob.B, chains, prob.D, generator=g, dtype=torch.float64) * 4 - 2
    r = prob.nuts(u0, chains=chains, warmup=warmup, samples=samples, seed=seed)
    assert torch.isfinite(r['draws']).all()
    assert (r['stepsize'] > 0).all()
    return r


def test_config3_series_outliers_posterior_matches_oracle():
    gold = np.load(os.path.join(GOLD, 'nuts_config_outliers.npz'))
    ds = oracle_batch(gold['freq'], [gold['Z']], mode='sample', outliers=True)
    prob = gpu_problem(ds)
    assert prob.D == 2 * prob.K + 9 + 2 * prob.Nf
    chains, warmup, samples = 16, 300, 600
    r = _nuts(prob, chains, warmup, samples, seed=11)
    assert r['n_divergent'].sum().item() <= 0.02 * chains * samples
    out = prob.split_outputs(prob.constrain(r['draws']))
    K, Nf = prob.K, prob.Nf
    cons = torch.cat([out['x'], out['Rinf'][..., None], out['induc'][..., None], out['sigma_res'][..., None],
                      out['alpha_prop'][..., None], out['alpha_re'][..., None], out['alpha_im'][..., None],
                      out['sigma_out']], dim=-1)[0].cpu().numpy()  # [chains, samples, K + 6 + Nf]
    assert list(gold['names'][[0, K, K + 6]]) == ['x', 'Rinf', 'sigma_out'] and cons.shape[-1] == len(gold['names'])
    rep = _compare(cons, gold, [np.arange(K), np.arange(K + 6, K + 6 + Nf)])
    print('config 3 |z| median / fraction beyond 3 / max:', rep)
    # the sampler finds the injected outliers: sigma_out at points 20, 45, 46 stands far above the rest
    so = cons[..., K + 6:].reshape(-1, Nf).mean(0)
    order = np.argsort(gold['freq'])[::-1]  # the model sorts by descending frequency
    flagged = set(np.argsort(so)[-3:])
    assert flagged == {int(np.where(order == i)[0][0]) for i in (20, 45, 46)}, flagged
    assert 0.3 < np.median(r['stepsize'].cpu().numpy()) / np.median(gold['stepsize']) < 3.0


def test_config5_series_parallel_pos_posterior_matches_oracle():
    from oracle import model_sp as osp
    gold = np.load(os.path.join(GOLD, 'nuts_config_sp.npz'))
    bf = np.logspace(6, -2, 81)
    ser = {'kernel': 'DRT', 'dist_type': 'series', 'basis_freq': bf}
    par = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'transmissive', 'dist_type': 'parallel', 'basis_freq': bf,
           'x_scale': 0.8}
    d = osp.prep_series_parallel(gold['freq'], gold['Z'], ser, par, mode='sample', nonneg=True)
    prob = gpu_problem_sp([d])
    assert prob.D == 2 * (81 + 81) + 12
    chains, warmup, samples = 16, 300, 600
    r = _nuts(prob, chains, warmup, samples, seed=12)
    assert r['n_divergent'].sum().item() <= 0.02 * chains * samples
    out = prob.split_outputs(prob.constrain(r['draws']))
    cons = torch.cat([out['xs'], out['xp'], out['Rinf'][..., None], out['induc'][..., None],
                      out['sigma_res'][..., None], out['alpha_prop'][..., None], out['alpha_re'][..., None],
                      out['alpha_im'][..., None]], dim=-1)[0].cpu().numpy()  # [chains, samples, Ks + Kp + 6]
    assert list(gold['names'][[0, 81, 162]]) == ['xs', 'xp', 'Rinf'] and cons.shape[-1] == len(gold['names'])
    rep = _compare(cons, gold, [np.arange(81), np.arange(81, 162)])
    print('config 5 |z| median / fraction beyond 3 / max:', rep)
    # trajectory lengths of the same order as the oracle's chains (the posterior of the real spectrum is well behaved:
    # ~550 leapfrogs per iteration, no saturation of the tree depth)
    leap = r['n_leapfrog'].sum().item() / (chains * (warmup + samples))
    assert 0.5 < leap / (gold['n_leapfrog'].mean() / gold['samples']) < 2.0
    assert 0.3 < np.median(r['stepsize'].cpu().numpy()) / np.median(gold['stepsize']) < 3.0
