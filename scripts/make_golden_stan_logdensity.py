"""Golden log-densities and gradients of the reference's Stan programs, evaluated MECHANICALLY from their source text
(bayes_drt/stan_model_files/*_modelcode.txt, read in place) by scripts/stan_subset_interpreter.py
-> tests/golden/stan_logdensity.npz.  The oracle's hand-derived log-posterior (oracle/model.py, oracle/model_sp.py) is
then checked against these values, which removes hand transcription from the parity chain of the Stan programs.
Run:  python scripts/make_golden_stan_logdensity.py   (needs /root/reference; build container only)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from oracle import model as omod, model_sp as osp  # noqa: E402  (data preparation only; pinned by stan_data.npz)
from stan_subset_interpreter import Program  # noqa: E402

SRC = '/root/reference/bayes_drt/stan_model_files'
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'spectra.npz'))
freq, Z = g['ZARC_uniform_0.25/freq'], g['ZARC_uniform_0.25/Z']
bf = np.logspace(6, -2, 41)  # smaller bases keep the fixture small
TP = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'transmissive', 'dist_type': 'parallel', 'basis_freq': bf}
BP = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'blocking', 'dist_type': 'parallel', 'basis_freq': bf}
DRT = {'kernel': 'DRT', 'basis_freq': bf}
fsub = freq[::2]
Zsub = Z[::2]


def series_data(d):
    Nf = d['Nf']
    dat = dict(N=Nf if d['outliers'] else 2 * Nf, K=d['K'], A=d['A'], Z=d['Z'], freq=d['freq'], L0=d['L0'], L1=d['L1'],
               L2=d['L2'], sigma_min=d['sigma_min'], ups_alpha=d['ups_alpha'], ups_beta=d['ups_beta'],
               induc_scale=d['induc_scale'])
    if d['outliers']:
        dat.update(sigma_out_lambda=d['sigma_out_lambda'], sigma_out_alpha=d['sigma_out_alpha'],
                   sigma_out_beta=d['sigma_out_beta'])
    else:
        dat.update(N_tilde=2 * Nf, A_tilde=d['A'], freq_tilde=d['freq'])
    return dat


def sp_data(d):
    Nf = d['Nf']
    dat = dict(N=2 * Nf, Ks=d['Ks'], As=d['As'], Z=d['Z'], freq=d['freq'], N_tilde=2 * Nf, As_tilde=d['As'],
               freq_tilde=d['freq'], L0s=d['Ls'][0], L1s=d['Ls'][1], L2s=d['Ls'][2], sigma_min=d['sigma_min'],
               ups_alpha=d['ups_alpha'], ups_beta=d['ups_beta'], induc_scale=d['induc_scale'],
               x_sum_invscale=d['x_sum_invscale'])
    if 'Kp2' in d:
        dat.update(Kp1=d['Kp'], Ap1=d['Ap'], Ap1_tilde=d['Ap'], L0p1=d['Lp'][0], L1p1=d['Lp'][1], L2p1=d['Lp'][2],
                   xp1_scale=d['xp_scale'], Kp2=d['Kp2'], Ap2=d['Ap2'], Ap2_tilde=d['Ap2'], L0p2=d['Lp2'][0],
                   L1p2=d['Lp2'][1], L2p2=d['Lp2'][2], xp2_scale=d['xp2_scale'])
    else:
        dat.update(Kp=d['Kp'], Ap=d['Ap'], Ap_tilde=d['Ap'], L0p=d['Lp'][0], L1p=d['Lp'][1], L2p=d['Lp'][2],
                   xp_scale=d['xp_scale'])
    return dat


CASES = {}
for mode in ('optimize', 'sample'):
    for nonneg in (False, True):
        for outl in (False, True):
            name = 'Series' + ('_pos' if nonneg else '') + ('_outliers' if outl else '')
            d = omod.prep_series(fsub, Zsub, basis_freq=bf, mode=mode, nonneg=nonneg, outliers=outl)
            CASES[f'{name}/{mode}'] = (name, series_data(d), dict(kind='series', mode=mode, nonneg=nonneg, outliers=outl))
    d = omod.prep_parallel(fsub, Zsub, TP, mode=mode)
    CASES[f'Parallel/{mode}'] = ('Parallel', series_data(d), dict(kind='parallel', mode=mode))
    for nonneg in (False, True):
        name = 'Series-Parallel' + ('_pos' if nonneg else '')
        d = osp.prep_series_parallel(fsub, Zsub, DRT, dict(TP, x_scale=0.8), mode=mode, nonneg=nonneg)
        CASES[f'{name}/{mode}'] = (name, sp_data(d), dict(kind='sp', mode=mode, nonneg=nonneg))
        name = 'Series-2Parallel' + ('_pos' if nonneg else '')
        d = osp.prep_series_2parallel(fsub, Zsub, DRT, BP, dict(TP, x_scale=0.8), mode=mode, nonneg=nonneg)
        CASES[f'{name}/{mode}'] = (name, sp_data(d), dict(kind='s2p', mode=mode, nonneg=nonneg))

out = {'freq': fsub, 'Z': Zsub, 'basis_freq': bf}
rng = np.random.RandomState(21)
for key, (name, dat, meta) in CASES.items():
    prog = Program(open(os.path.join(SRC, name + '_modelcode.txt')).read())
    D = prog.n_unconstrained(dat)
    U = rng.uniform(-1.0, 1.0, (3, D))
    U[:, 0] -= 4.0  # Rinf = 100 exp(u): keep the offset of the size of the data
    if meta['kind'] in ('sp', 's2p') and not meta['nonneg']:
        U[2, 2:2 + dat['Ks']] = -3.0  # xs free in sign: a point with sum(xs) + sum(xp_raw) < 0 -> rejected (x_sum_raw >= 0)
    out[key + '/U'] = U
    for jac in (False, True):
        lps, grads = [], []
        for i in range(3):
            u = torch.tensor(U[i], requires_grad=True)
            lp = prog.log_prob(dat, u, jac)
            if torch.isfinite(lp):
                lp.backward()
                grads.append(u.grad.numpy().copy())
            else:
                grads.append(np.full(D, np.nan))
            lps.append(float(lp))
        out[key + f'/lp_jac{int(jac)}'] = np.array(lps)
        out[key + f'/grad_jac{int(jac)}'] = np.array(grads)
    env = {}
    prog.log_prob(dat, torch.tensor(U[0]), False, env_out=env)  # what Stan reports for the first point
    for nm in ('x', 'xs', 'xp', 'xp1', 'xp2', 'Rinf', 'induc', 'sigma_res', 'alpha_prop', 'alpha_re', 'alpha_im',
               'sigma_tot', 'sigma_out', 'Z_hat'):
        if nm in env:
            out[key + '/tp/' + nm] = env[nm]
    print(key, 'D', D, 'lp', out[key + '/lp_jac0'])
dst = os.path.join(ROOT, 'tests', 'golden', 'stan_logdensity.npz')
np.savez_compressed(dst, **out)
print('wrote', dst, os.path.getsize(dst), 'bytes')
