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
