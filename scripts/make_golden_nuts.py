"""Long oracle NUTS run -> tests/golden/nuts_<name>.npz (posterior summaries used by the statistical parity test of the
CUDA sampler).  Build container only; takes a few minutes on 8 cores.

    python scripts/make_golden_nuts.py [name] [chains] [warmup] [samples]
"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from oracle import model as omod, nuts  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'ZARC-RL_uniform_0.25'
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 8
warmup = int(sys.argv[3]) if len(sys.argv) > 3 else 300
samples = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
g = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'spectra.npz'))
d = omod.prep_series(g[name + '/freq'], g[name + '/Z'], mode='sample')
D = omod.n_params(d)


def run(c):
    rng = np.random.RandomState(1000 + c)
    u0 = rng.uniform(-2, 2, D)
    return nuts.sample_chain(lambda u: omod.logpost(u, d, jacobian=True), u0, warmup=warmup, samples=samples, seed=c)


if __name__ == '__main__':
    t = time.time()
    with Pool(min(chains, os.cpu_count())) as p:
        res = p.map(run, range(chains))
    draws = np.stack([r['draws'] for r in res])  # [chains, samples, D]
    K = d['K']
    # constrained quantities the reference reports: x, Rinf, induc, sigma_res, alpha_*
    cons = np.empty((chains, samples, K + 6))
    for c in range(chains):
        for s in range(samples):
            o = omod.constrain(draws[c, s], d)
            cons[c, s, :K] = o['x']
            cons[c, s, K:] = [o['Rinf'], o['induc'], o['sigma_res'], o['alpha_prop'], o['alpha_re'], o['alpha_im']]
    flat = cons.reshape(-1, K + 6)
    ess = np.array([nuts.ess_bulk(cons[:, :, i]) for i in range(K + 6)])
    out = dict(mean=flat.mean(0), sd=flat.std(0, ddof=1), q025=np.percentile(flat, 2.5, axis=0),
               q975=np.percentile(flat, 97.5, axis=0), q50=np.percentile(flat, 50, axis=0), ess=ess,
               stepsize=np.array([r['stepsize'] for r in res]), n_leapfrog=np.array([r['n_leapfrog'] for r in res]),
               n_divergent=np.array([r['n_divergent'] for r in res]), n_maxdepth=np.array([r['n_maxdepth'] for r in res]),
               chains=chains, warmup=warmup, samples=samples, Z_scale=d['Z_scale'], wall_s=time.time() - t,
               chain_means=cons.mean(1),
               mcse_mean=np.array([nuts.mcse_mean(cons[:, :, i]) for i in range(K + 6)]),
               mcse_q025=np.array([nuts.mcse_quantile(cons[:, :, i], 0.025) for i in range(K + 6)]),
               mcse_q975=np.array([nuts.mcse_quantile(cons[:, :, i], 0.975) for i in range(K + 6)]))
    dst = os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', f'nuts_{name}.npz')
    np.savez_compressed(dst, **out)
    print('wrote', dst, 'wall %.0f s' % (time.time() - t), 'min ESS', ess.min(), 'stepsizes', out['stepsize'],
          'leapfrogs/iter', out['n_leapfrog'] / samples, 'div', out['n_divergent'])
