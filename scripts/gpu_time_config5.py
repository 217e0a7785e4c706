"""BASELINE config 5 timing: Series-Parallel_pos (DRT + TP-DDT, Ks = Kp = 81, D = 336), every spectrum on its own
frequency grid freq_b = 10**(6 - delta_b - arange(81)/10), batched HMC (4 chains x (200 + 200))."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from bayes_drt_b200 import capi
from helpers import sp_spectrum
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rng = np.random.RandomState(0)
bf = np.logspace(6, -2, 81)
tau = torch.tensor(1 / (2 * np.pi * bf))
eps = 1 / np.mean(np.diff(np.log(tau.numpy())))
fr = np.stack([10 ** (6 - rng.uniform(0, 1) - np.arange(81) / 10) for _ in range(B)])
Z = np.stack([sp_spectrum(f, seed=b, td=10 ** rng.uniform(-1.5, 0), tau0=10 ** rng.uniform(-4, -2)) for b, f in enumerate(fr)])
zs = np.std(np.abs(Z), axis=1)
Zst = torch.tensor(np.concatenate((Z.real, Z.imag), axis=1) / zs[:, None])
fb = torch.tensor(fr)
capi.build_A(fb[:2], tau, eps); torch.cuda.synchronize()  # context / library warm-up
t0 = time.time()
As = capi.build_A(fb, tau, eps)
Ap = capi.build_A(fb, tau, eps, kernel='DDT', dist_type='parallel', symmetry='planar', bc='transmissive')
torch.cuda.synchronize(); t_mat = time.time() - t0
Lb = [capi.build_L(torch.tensor(bf), tau, eps, o) for o in range(3)]
L = torch.stack([Lb[0], Lb[1], 0.75 * Lb[2]])
prob = capi.SeriesProblem(torch.cat(As, dim=1), Zst, fb, L, nonneg=True, ups_alpha=1.0, ups_beta=0.1,
                          Ap=torch.cat(Ap, dim=1), Lp=L, x_sum_invscale=1.0, xp_scale=0.8)
g = torch.Generator().manual_seed(0)
u0 = (torch.rand(B, chains, prob.D, generator=g, dtype=torch.float64) * 4 - 2).cuda()
prob.nuts(u0, chains=chains, warmup=3, samples=1, keep_draws=False)  # warm-up launch
torch.cuda.synchronize(); t = time.time()
r = prob.nuts(u0, chains=chains, warmup=200, samples=200, keep_draws=False)
torch.cuda.synchronize(); dt = time.time() - t
ng = r['n_leapfrog'].sum().item()
print(f'config5: B={B} per-spectrum grids, D={prob.D}, {chains} chains x (200+200): matrices {t_mat*1e3:.1f} ms ({2*B/t_mat:.0f} A builds/s), '
      f'NUTS {dt:.2f} s -> {B/dt:.2f} inversions/s, {ng/dt:.3e} grads/s, leap/iter {ng/(B*chains*400):.0f}, '
      f'div {r["n_divergent"].sum().item()}, maxdepth frac {r["n_maxdepth"].float().mean().item()/200:.3f}')
