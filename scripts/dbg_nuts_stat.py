"""GPU diagnostic: z-scores of the NUTS posterior summaries vs the oracle golden, per parameter group, over seeds / layouts."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from helpers import GOLD, gpu_problem, load_spectrum, oracle_batch
from oracle.nuts import ess_bulk, mcse_mean, mcse_quantile
NAME = 'ZARC-RL_uniform_0.25'
gold = np.load(os.path.join(GOLD, f'nuts_{NAME}.npz'))
freq, Z = load_spectrum(NAME)
ds = oracle_batch(freq, [Z], mode='sample')
for dense in ('0', '1'):
    os.environ['BDRT_FORCE_DENSE'] = dense
    prob = gpu_problem(ds)
    K = prob.K
    for seed in (7, 8, 9, 10):
        chains, warmup, samples = 16, 300, 500
        g = torch.Generator().manual_seed(seed)
        u0 = torch.rand(prob.B, chains, prob.D, generator=g, dtype=torch.float64) * 4 - 2
        r = prob.nuts(u0, chains=chains, warmup=warmup, samples=samples, seed=seed)
        out = prob.split_outputs(prob.constrain(r['draws']))
        cons = torch.cat([out['x']] + [out[k][..., None] for k in ('Rinf', 'induc', 'sigma_res', 'alpha_prop', 'alpha_re', 'alpha_im')], dim=-1)[0].cpu().numpy()
        flat = cons.reshape(-1, K + 6)
        mean, sd = flat.mean(0), flat.std(0, ddof=1)
        q025, q975 = np.percentile(flat, 2.5, axis=0), np.percentile(flat, 97.5, axis=0)
        ess = np.array([ess_bulk(cons[:, :, i]) for i in range(K + 6)])
        ses = {'mean': np.hypot([mcse_mean(cons[:, :, i]) for i in range(K + 6)], gold['mcse_mean']),
               'q025': np.hypot([mcse_quantile(cons[:, :, i], 0.025) for i in range(K + 6)], gold['mcse_q025']),
               'q975': np.hypot([mcse_quantile(cons[:, :, i], 0.975) for i in range(K + 6)], gold['mcse_q975'])}
        scale = np.abs(gold['mean'][:K]).max()
        msg = []
        for nm, a, b, f in (('mean', mean, gold['mean'], 1.0), ('q025', q025, gold['q025'], 2.67), ('q975', q975, gold['q975'], 2.67)):
            z = np.abs(a - b) / ses[nm]
            big = np.abs(a - b) > 2e-3 * np.r_[np.full(K, scale), np.abs(b[K:]) + 1e-12]
            zz = np.where(big, z, 0)
            i = int(np.argmax(zz))
            msg.append(f'{nm}: max z {zz.max():.1f} @ {i} (ess {ess[i]:.0f}, a {a[i]:.4g} b {b[i]:.4g}) n>3: {(zz>3).sum()}')
        cm = cons.mean(axis=1)  # chain means
        rhat_like = (cm.std(0, ddof=1) / (sd + 1e-300)).max()
        print(f'dense={dense} seed={seed} step med {r["stepsize"].median().item():.4f} [{r["stepsize"].min().item():.4f},{r["stepsize"].max().item():.4f}] div {r["n_divergent"].sum().item()} maxd {r["n_maxdepth"].sum().item()} min ess {ess.min():.0f} chain-mean-spread/sd max {rhat_like:.2f} | ' + ' | '.join(msg), flush=True)
