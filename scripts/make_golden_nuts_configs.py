"""Long oracle NUTS runs for BASELINE.json configs 3 and 5 -> tests/golden/nuts_<case>.npz (posterior summaries with their
own Monte-Carlo standard errors, used by the 3-MCSE parity tests of the CUDA sampler in tests/test_gpu_nuts_configs.py).
Build container only (reads /root/reference/data for config 5); minutes to an hour on 8 cores.

    python scripts/make_golden_nuts_configs.py outliers [chains] [warmup] [samples]
        config 3: Stan program Series_outliers (inversion.py:1218-1221 with outliers=True; Series_outliers_modelcode.txt),
        spectrum ZARC_uniform_0.25 with three injected outliers (recipe below; the spectrum itself is stored in the file)
    python scripts/make_golden_nuts_configs.py sp [chains] [warmup] [samples]
        config 5: Stan program Series-Parallel_pos on data/simulated/Z_DRT-2-TpDDT_uniform_0.25.csv with the paper's
        distributions (code_EchemActa/Run fits.ipynb cell 20: DRT + transmissive planar DDT in parallel form, both on
        basis_freq = logspace(6, -2, 81), x_scale 0.8)
"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from oracle import model as omod, model_sp as osp, nuts  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else 'outliers'
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 16
warmup = int(sys.argv[3]) if len(sys.argv) > 3 else 300
samples = int(sys.argv[4]) if len(sys.argv) > 4 else 2000


def outlier_spectrum():
    """ZARC_uniform_0.25 (tests/golden/spectra.npz) with three gross errors, the kind Tutorial 3 treats (a few isolated
    points off by many sigma): +8 % / -6 % / +5 % of |Z| range on points 20, 45, 46 (real and imaginary parts)."""
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'spectra.npz'))
    freq, Z = g['ZARC_uniform_0.25/freq'].copy(), g['ZARC_uniform_0.25/Z'].copy()
    rng_ = np.abs(Z).max() - np.abs(Z).min()
    Z[20] += 0.08 * rng_ * (1 + 1j)
    Z[45] -= 0.06 * rng_ * (1 - 0.5j)
    Z[46] += 0.05 * rng_ * (0.3 + 1j)
    return freq, Z


def sp_setup():
    import pandas as pd
    df = pd.read_csv('/root/reference/data/simulated/Z_DRT-2-TpDDT_uniform_0.25.csv')
    freq = df['Freq'].values
    Z = df['Zreal'].values + 1j * df['Zimag'].values
    bf = np.logspace(6, -2, 81)
    ser = {'kernel': 'DRT', 'dist_type': 'series', 'basis_freq': bf}
    par = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'transmissive', 'dist_type': 'parallel', 'basis_freq': bf,
           'x_scale': 0.8}
    return freq, Z, ser, par


if case == 'outliers':
    freq, Z = outlier_spectrum()
    d = omod.prep_series(freq, Z, mode='sample', outliers=True)
    D = omod.n_params(d)
    lpf = lambda u: omod.logpost(u, d, jacobian=True)  # noqa: E731

    def cons_of(u):
        o = omod.constrain(u, d)
        return np.r_[o['x'], o['Rinf'], o['induc'], o['sigma_res'], o['alpha_prop'], o['alpha_re'], o['alpha_im'],
                     o['sigma_out']]
    names = ['x'] * d['K'] + ['Rinf', 'induc', 'sigma_res', 'alpha_prop', 'alpha_re', 'alpha_im'] + ['sigma_out'] * d['Nf']
elif case == 'sp':
    freq, Z, ser, par = sp_setup()
    d = osp.prep_series_parallel(freq, Z, ser, par, mode='sample', nonneg=True)
    D = osp.n_params(d)
    lpf = lambda u: osp.logpost(u, d, jacobian=True)  # noqa: E731

    def cons_of(u):
        o = osp.constrain(u, d)
        return np.r_[o['xs'], o['xp'], o['Rinf'], o['induc'], o['sigma_res'], o['alpha_prop'], o['alpha_re'],
                     o['alpha_im']]
    names = ['xs'] * d['Ks'] + ['xp'] * d['Kp'] + ['Rinf', 'induc', 'sigma_res', 'alpha_prop', 'alpha_re', 'alpha_im']
else:
    raise SystemExit('case must be outliers or sp')


def run(c):
    rng = np.random.RandomState(1000 + c)
    for attempt in range(100):  # Stan retries random inits until lp and gradient are finite
        u0 = rng.uniform(-2, 2, D)
        lp, g = lpf(u0)
        if np.isfinite(lp) and np.all(np.isfinite(g)):
            break
    return nuts.sample_chain(lpf, u0, warmup=warmup, samples=samples, seed=c)


def rhat(x):
    """split-chain potential scale reduction of draws x [chains, n]"""
    c, n = x.shape
    half = n // 2
    z = np.concatenate((x[:, :half], x[:, half:2 * half]), axis=0)
    W = z.var(axis=1, ddof=1).mean()
    Bv = z.mean(axis=1).var(ddof=1)
    return np.sqrt(((half - 1) / half * W + Bv) / W)


if __name__ == '__main__':
    t = time.time()
    with Pool(min(chains, os.cpu_count())) as p:
        res = p.map(run, range(chains))
    draws = np.stack([r['draws'] for r in res])  # [chains, samples, D]
    P = len(names)
    cons = np.empty((chains, samples, P))
    for c in range(chains):
        for s in range(samples):
            cons[c, s] = cons_of(draws[c, s])
    flat = cons.reshape(-1, P)
    out = dict(case=case, names=np.array(names), freq=freq, Z=Z,
               mean=flat.mean(0), sd=flat.std(0, ddof=1), q025=np.percentile(flat, 2.5, axis=0),
               q975=np.percentile(flat, 97.5, axis=0), q50=np.percentile(flat, 50, axis=0),
               ess=np.array([nuts.ess_bulk(cons[:, :, i]) for i in range(P)]),
               rhat=np.array([rhat(cons[:, :, i]) for i in range(P)]),
               stepsize=np.array([r['stepsize'] for r in res]), n_leapfrog=np.array([r['n_leapfrog'] for r in res]),
               n_divergent=np.array([r['n_divergent'] for r in res]), n_maxdepth=np.array([r['n_maxdepth'] for r in res]),
               chains=chains, warmup=warmup, samples=samples, Z_scale=d['Z_scale'], wall_s=time.time() - t,
               chain_means=cons.mean(1),
               mcse_mean=np.array([nuts.mcse_mean(cons[:, :, i]) for i in range(P)]),
               mcse_q025=np.array([nuts.mcse_quantile(cons[:, :, i], 0.025) for i in range(P)]),
               mcse_q975=np.array([nuts.mcse_quantile(cons[:, :, i], 0.975) for i in range(P)]))
    dst = os.path.join(ROOT, 'tests', 'golden', f'nuts_config_{case}.npz')
    np.savez_compressed(dst, **out)
    print('wrote', dst, 'wall %.0f s' % (time.time() - t), 'min ESS', out['ess'].min(), 'max rhat', out['rhat'].max(),
          'stepsizes', out['stepsize'], 'leapfrogs/iter', out['n_leapfrog'] / samples, 'div', out['n_divergent'],
          'maxdepth', out['n_maxdepth'])
