"""GPU: per-spectrum frequency grids (BASELINE config 5): every spectrum brings its own kernel matrices; a CTA takes one
spectrum at a time and its slots run that spectrum's chains.  Checked against the oracle and against the same spectra
run one by one as shared-grid problems."""
import numpy as np
import pytest
import torch

from helpers import sp_dists, sp_spectrum
from oracle import model as omod, model_sp as osp

pytestmark = pytest.mark.gpu


def _series_problem(freqs, Zs, bf, mode, **kw):
    """per-spectrum SeriesProblem + the oracle dicts (matrices of each grid built by the CUDA path in one batched call)"""
    from bayes_drt_b200 import capi
    ds = [omod.prep_series(f, Z, basis_freq=bf, mode=mode, **kw) for f, Z in zip(freqs, Zs)]
    tau, eps = ds[0]['tau'], ds[0]['epsilon']
    fb = torch.tensor(np.stack([d['freq'] for d in ds]))
    A_re, A_im = capi.build_A(fb, torch.tensor(tau), eps)  # [B, Nf, K]
    c = omod.MODE_CONSTANTS[mode]
    L = torch.stack([c[f'l{o}'] * capi.build_L(torch.tensor(1 / (2 * np.pi * tau)), torch.tensor(tau), eps, o)
                     for o in range(3)])
    prob = capi.SeriesProblem(torch.cat((A_re, A_im), dim=1), torch.tensor(np.stack([d['Z'] for d in ds])), fb, L,
                              nonneg=ds[0]['pos'], outliers=ds[0]['outliers'], ups_alpha=ds[0]['ups_alpha'],
                              ups_beta=ds[0]['ups_beta'])
    assert prob.per_spectrum_grid
    return prob, ds


def _one(prob_all, ds, b):
    """spectrum b alone, as a shared-grid problem with the same matrices"""
    from bayes_drt_b200 import capi
    d = ds[b]
    return capi.SeriesProblem(prob_all.A[b], prob_all.Z[b:b + 1], prob_all.freq[b], prob_all.L, nonneg=d['pos'],
                              outliers=d['outliers'], ups_alpha=d['ups_alpha'], ups_beta=d['ups_beta'])


@pytest.mark.parametrize('loguniform', [True, False])
def test_series_per_spectrum_grids(loguniform, monkeypatch):
    # per-spectrum grids run the L-BFGS driver in warp mode (eight spectra with their own tables per CTA); the shared-grid
    # single runs they are compared with bit for bit below are put on the same engine (their default is the cooperative
    # one, whose products round differently)
    monkeypatch.setenv('BDRT_WARP', '1')
    rng = np.random.RandomState(2)
    bf = np.logspace(6, -2, 81)
    freqs, Zs = [], []
    for b in range(3):
        f = 10 ** (6 - rng.uniform(0, 1) - np.arange(71) / 10)  # offset grids, same spacing -> still Toeplitz
        if not loguniform:
            f = f * np.exp(rng.uniform(-0.03, 0.03, len(f)))    # jitter -> dense resident operands
        freqs.append(f)
        Zs.append(sp_spectrum(f, seed=b, Rd=0.0, tau0=10 ** rng.uniform(-4, -1)))
    prob, ds = _series_problem(freqs, Zs, bf, 'optimize')
    u = rng.uniform(-1, 1, (7, prob.D))
    spec = np.array([0, 1, 2, 2, 1, 0, 1])
    lp, grad = prob.logpost_grad(torch.tensor(u), spec=spec)
    for c in range(7):
        lo, go = omod.logpost(u[c], ds[spec[c]])
        assert abs(lp[c].item() - lo) <= 1e-11 * abs(lo)
        assert np.max(np.abs(grad[c].cpu().numpy() - go)) <= 1e-9 * np.max(np.abs(go))
    u0 = torch.tensor(rng.uniform(-2, 2, (3, prob.D)))
    r = prob.map_lbfgs(u0, max_iter=150)
    assert (r['lp'] > prob.logpost_grad(u0, spec=np.arange(3))[0]).all() and (r['iters'] == 150).all()
    for b in range(3):
        rb = _one(prob, ds, b).map_lbfgs(u0[b:b + 1], max_iter=150)
        assert torch.allclose(rb['u'][0], r['u'][b], rtol=0, atol=1e-9 * r['u'][b].abs().max().item())
        assert rb['n_eval'][0].item() == r['n_eval'][b].item()
    p = prob.map_newton(r['u'], max_iter=3)
    assert (p['lp'] >= r['lp'] - 1e-9).all()


def test_series_per_spectrum_nuts_matches_single_runs():
    rng = np.random.RandomState(3)
    bf = np.logspace(6, -2, 81)
    freqs = [10 ** (6 - d - np.arange(71) / 10) for d in (0.0, 0.37, 0.81)]
    Zs = [sp_spectrum(f, seed=b, Rd=0.0) for b, f in enumerate(freqs)]
    prob, ds = _series_problem(freqs, Zs, bf, 'sample')
    g = torch.Generator().manual_seed(1)
    for chains in (2, 10):  # fewer and more chains than slots
        u0 = torch.rand(3, chains, prob.D, generator=g, dtype=torch.float64) * 4 - 2
        kw = dict(chains=chains, warmup=25, samples=6, seed=11)
        a = prob.nuts(u0, **kw)
        assert torch.isfinite(a['draws']).all() and (a['stepsize'] > 0).all()
        b1 = _one(prob, ds, 1).nuts(u0[1:2], spectrum_offset=1, **kw)
        assert torch.allclose(b1['draws'][0], a['draws'][1], rtol=0, atol=1e-9)
        assert torch.equal(b1['n_leapfrog'][0], a['n_leapfrog'][1])


def test_config5_series_parallel_per_spectrum_grids():
    """BASELINE config 5: DRT + TP-DDT (Series-Parallel_pos, Ks = Kp = 81, D = 336), each spectrum on its own grid
    freq_b = 10**(6 - delta_b - arange(81)/10), batched HMC."""
    from bayes_drt_b200 import capi
    rng = np.random.RandomState(5)
    bf = np.logspace(6, -2, 81)
    ser, par = sp_dists(bf, bf)
    ds, fr = [], []
    for b in range(3):
        f = 10 ** (6 - rng.uniform(0, 1) - np.arange(81) / 10)
        ds.append(osp.prep_series_parallel(f, sp_spectrum(f, seed=b, td=0.1 * (b + 1)), ser, par, mode='sample'))
        fr.append(ds[-1]['freq'])
    d0 = ds[0]
    fb = torch.tensor(np.stack(fr))
    tau = torch.tensor(d0['tau_s'])
    As = capi.build_A(fb, tau, d0['eps_s'])
    Ap = capi.build_A(fb, tau, d0['eps_p'], kernel='DDT', dist_type='parallel', symmetry='planar', bc='transmissive')
    c = osp.MODE_CONSTANTS_SP['sample']
    Lb = [capi.build_L(torch.tensor(bf), tau, d0['eps_s'], o) for o in range(3)]
    prob = capi.SeriesProblem(torch.cat(As, dim=1), torch.tensor(np.stack([d['Z'] for d in ds])), fb,
                              torch.stack([c['ls'][o] * Lb[o] for o in range(3)]), nonneg=True, ups_alpha=1.0, ups_beta=0.1,
                              Ap=torch.cat(Ap, dim=1), Lp=torch.stack([c['lp'][o] * Lb[o] for o in range(3)]),
                              x_sum_invscale=1.0, xp_scale=0.8)
    assert prob.D == 336 and prob.per_spectrum_grid
    u = rng.uniform(-1, 1, (5, prob.D))
    spec = np.array([2, 0, 1, 1, 2])
    lp, grad = prob.logpost_grad(torch.tensor(u), spec=spec, jacobian=True)
    for k in range(5):
        lo, go = osp.logpost(u[k], ds[spec[k]], jacobian=True)
        assert abs(lp[k].item() - lo) <= 1e-11 * abs(lo)
        assert np.max(np.abs(grad[k].cpu().numpy() - go)) <= 1e-9 * np.max(np.abs(go))
    g = torch.Generator().manual_seed(2)
    u0 = torch.rand(3, 4, prob.D, generator=g, dtype=torch.float64) * 4 - 2
    a = prob.nuts(u0, chains=4, warmup=40, samples=10, seed=3)
    assert torch.isfinite(a['draws']).all() and (a['accept'] > 0.05).all()
    out = prob.split_outputs(prob.constrain(a['draws'].reshape(-1, prob.D),
                                            spec=torch.arange(3, dtype=torch.int32).repeat_interleave(40)))
    assert (out['xs'] >= 0).all() and (out['xp'] >= 0).all() and torch.isfinite(out['sigma_tot']).all()
