"""Shared test helpers: build a SeriesProblem on the GPU next to the oracle's data dicts for the same spectra."""
import os

import numpy as np
import torch

from oracle import model as omod

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def load_spectrum(name):
    g = np.load(os.path.join(GOLD, 'spectra.npz'))
    return g[name + '/freq'], g[name + '/Z']


def oracle_batch(freq, Zs, basis_freq=None, **kw):
    """list of oracle data dicts (one per spectrum), sharing the A matrices."""
    d0 = omod.prep_series(freq, Zs[0], basis_freq=basis_freq, **kw)
    Nf = d0['Nf']
    out = [d0]
    for Z in Zs[1:]:
        out.append(omod.prep_series(freq, Z, basis_freq=basis_freq, A_re=d0['A'][:Nf], A_im=d0['A'][Nf:], **kw))
    return out


def gpu_problem(ds):
    """capi.SeriesProblem with kernel matrices built by the CUDA path (not copied from the oracle)."""
    from bayes_drt_b200 import capi
    d0 = ds[0]
    tau, eps = d0['tau'], d0['epsilon']
    A_re, A_im = capi.build_A(torch.tensor(d0['freq']), torch.tensor(tau), eps)
    bf = torch.tensor(1 / (2 * np.pi * tau))
    c = omod.MODE_CONSTANTS['optimize' if d0['ups_alpha'] == 0.05 else 'sample']
    L = torch.stack([c[f'l{o}'] * capi.build_L(bf, torch.tensor(tau), eps, o) for o in (0, 1, 2)])
    return capi.SeriesProblem(torch.cat((A_re, A_im)), torch.tensor(np.stack([d['Z'] for d in ds])),
                              torch.tensor(d0['freq']), L, nonneg=d0['pos'], outliers=d0['outliers'],
                              sigma_min=d0['sigma_min'], ups_alpha=d0['ups_alpha'], ups_beta=d0['ups_beta'],
                              induc_scale=d0['induc_scale'], sigma_out_lambda=d0['sigma_out_lambda'],
                              sigma_out_alpha=d0['sigma_out_alpha'], sigma_out_beta=d0['sigma_out_beta'])


def sp_dists(bf_s=None, bf_p=None, x_scale=0.8):
    """The paper's two-distribution setup (code_EchemActa/Run fits.ipynb cell 20): DRT in series with a transmissive
    planar DDT in parallel form."""
    ser = {'kernel': 'DRT', 'dist_type': 'series'}
    par = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'transmissive', 'dist_type': 'parallel', 'x_scale': x_scale}
    if bf_s is not None:
        ser['basis_freq'] = bf_s
    if bf_p is not None:
        par['basis_freq'] = bf_p
    return ser, par


def sp_spectrum(freq, seed=0, Rinf=0.5, R1=1.0, tau0=1e-3, n=0.8, Rd=0.7, td=0.3, noise=0.002):
    """ZARC in series with a finite-length (transmissive) Warburg element."""
    rng = np.random.RandomState(seed)
    w = 2 * np.pi * np.asarray(freq)
    Zd = Rd * np.tanh(np.sqrt(1j * w * td)) / np.sqrt(1j * w * td)
    Z = Rinf + R1 / (1 + (1j * w * tau0) ** n) + Zd
    return Z + noise * (rng.standard_normal(len(w)) + 1j * rng.standard_normal(len(w)))


def gpu_problem_sp(ds):
    """capi.SeriesProblem for a list of oracle.model_sp data dicts sharing one grid; matrices built by the CUDA path."""
    from bayes_drt_b200 import capi
    from oracle.model_sp import MODE_CONSTANTS_SP
    d0 = ds[0]
    f = torch.tensor(d0['freq'])
    As_re, As_im = capi.build_A(f, torch.tensor(d0['tau_s']), d0['eps_s'])
    Ap_re, Ap_im = capi.build_A(f, torch.tensor(d0['tau_p']), d0['eps_p'], kernel='DDT', dist_type='parallel',
                                symmetry='planar', bc='transmissive')
    c = MODE_CONSTANTS_SP['optimize' if d0['ups_alpha'] == 0.05 else 'sample']
    Ls = torch.stack([c['ls'][o] * capi.build_L(torch.tensor(1 / (2 * np.pi * d0['tau_s'])), torch.tensor(d0['tau_s']),
                                                 d0['eps_s'], o) for o in range(3)])
    Lp = torch.stack([c['lp'][o] * capi.build_L(torch.tensor(1 / (2 * np.pi * d0['tau_p'])), torch.tensor(d0['tau_p']),
                                                 d0['eps_p'], o) for o in range(3)])
    return capi.SeriesProblem(torch.cat((As_re, As_im)), torch.tensor(np.stack([d['Z'] for d in ds])), f, Ls,
                              nonneg=d0['pos'], sigma_min=d0['sigma_min'], ups_alpha=d0['ups_alpha'],
                              ups_beta=d0['ups_beta'], induc_scale=d0['induc_scale'], Ap=torch.cat((Ap_re, Ap_im)),
                              Lp=Lp, x_sum_invscale=d0['x_sum_invscale'], xp_scale=d0['xp_scale'])
