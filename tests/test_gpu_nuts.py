"""GPU: batched on-device NUTS against a long run of the oracle's restatement of Stan's sampler
(tests/golden/nuts_*.npz from scripts/make_golden_nuts.py: 16 chains x (300 + 2000)): posterior means and 95 % interval end points of every DRT
coefficient and of R_inf / inductance / error-model parameters within 3 Monte-Carlo standard errors (north star).
Random streams cannot match (Philox vs numpy), so parity is statistical."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLD, gpu_problem, load_spectrum, oracle_batch

pytestmark = pytest.mark.gpu
NAME = 'ZARC-RL_uniform_0.25'


def _ess(x):
    from oracle.nuts import ess_bulk
    return ess_bulk(x)


def _run(prob, chains, warmup, samples, seed, spectrum_offset=0):
    g = torch.Generator().manual_seed(seed)
    u0 = torch.rand(prob.B, chains, prob.D, generator=g, dtype=torch.float64) * 4 - 2
    return prob.nuts(u0, chains=chains, warmup=warmup, samples=samples, seed=seed, spectrum_offset=spectrum_offset)


def test_nuts_posterior_matches_oracle():
    gold = np.load(os.path.join(GOLD, f'nuts_{NAME}.npz'))
    freq, Z = load_spectrum(NAME)
    ds = oracle_batch(freq, [Z], mode='sample')
    prob = gpu_problem(ds)
    chains, warmup, samples = 16, 300, 500
    r = _run(prob, chains, warmup, samples, seed=7)
    assert torch.isfinite(r['draws']).all()
    assert (r['stepsize'] > 0).all() and (r['n_divergent'].sum().item() <= 0.02 * chains * samples)
    out = prob.split_outputs(prob.constrain(r['draws']))
    K = prob.K
    cons = torch.cat([out['x'], out['Rinf'][..., None], out['induc'][..., None], out['sigma_res'][..., None],
                      out['alpha_prop'][..., None], out['alpha_re'][..., None], out['alpha_im'][..., None]],
                     dim=-1)[0].cpu().numpy()  # [chains, samples, K+6]
    from oracle.nuts import mcse_mean, mcse_quantile
    flat = cons.reshape(-1, K + 6)
    mean = flat.mean(0)
    q025, q975 = np.percentile(flat, 2.5, axis=0), np.percentile(flat, 97.5, axis=0)
    # Monte-Carlo standard errors of both sides (Vehtari et al. 2021): sd / sqrt(ESS) for the mean; for a quantile the ESS
    # of the indicator I(x <= q) gives a Beta interval in probability that is mapped back through the empirical
    # quantile function.  The golden file (16 oracle chains x 2000 draws) carries its own MCSEs.
    se_mean = np.hypot([mcse_mean(cons[:, :, i]) for i in range(K + 6)], gold['mcse_mean'])
    se_lo = np.hypot([mcse_quantile(cons[:, :, i], 0.025) for i in range(K + 6)], gold['mcse_q025'])
    se_hi = np.hypot([mcse_quantile(cons[:, :, i], 0.975) for i in range(K + 6)], gold['mcse_q975'])
    z_mean = np.abs(mean - gold['mean']) / se_mean
    z_lo = np.abs(q025 - gold['q025']) / se_lo
    z_hi = np.abs(q975 - gold['q975']) / se_hi
    # north star: within 3 MCSE.  With 3 x 107 comparisons a few excursions beyond 3 are expected by chance (and the
    # MCSEs are themselves estimates), so: at most 3 % beyond 3 and none beyond 6; coefficients far in the tails of the
    # DRT are ~0 with tiny sd and are judged on the scale of the peak as well
    scale = np.abs(gold['mean'][:K]).max()
    for z, a, b in ((z_mean, mean, gold['mean']), (z_lo, q025, gold['q025']), (z_hi, q975, gold['q975'])):
        viol = (z > 3.0) & (np.abs(a - b) > 2e-3 * np.r_[np.full(K, scale), np.abs(b[K:]) + 1e-12])
        assert viol.mean() <= 0.03, (np.where(viol)[0], z[viol])
        assert np.all(z[viol] < 6.0) if viol.any() else True, (np.where(viol)[0], z[viol])
    # step sizes and tree depths of the same order as the oracle's chains
    assert 0.3 < np.median(r['stepsize'].cpu().numpy()) / np.median(gold['stepsize']) < 3.0


def test_nuts_deterministic_and_shard_independent(resident_A):
    from bayes_drt_b200 import synth
    freq, Z, _ = synth.make_spectra(6, seed=4)
    _, bf = synth.bench_grid()
    ds = oracle_batch(freq.numpy(), list(Z.numpy()), basis_freq=bf.numpy(), mode='sample')
    prob = gpu_problem(ds)
    g = torch.Generator().manual_seed(0)
    u0 = torch.rand(6, 2, prob.D, generator=g, dtype=torch.float64) * 4 - 2
    kw = dict(chains=2, warmup=30, samples=10, seed=99)
    a = prob.nuts(u0, **kw)
    b = prob.nuts(u0, **kw)
    assert torch.equal(a['draws'], b['draws'])
    # spectra 2..3 as their own batch with spectrum_offset=2 (what another GPU's shard would run)
    sub = gpu_problem(ds[2:4])
    c = sub.nuts(u0[2:4], spectrum_offset=2, **kw)
    assert torch.equal(a['draws'][2:4], c['draws'])
    assert torch.equal(a['n_leapfrog'][2:4], c['n_leapfrog'])
    assert torch.isfinite(a['draws']).all()
    assert (a["accept"] > 0.01).all() and (a["stepsize"] > 0).all()  # warmup=30 is too short to adapt well
