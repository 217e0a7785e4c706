"""Dump per-spectrum L-BFGS iteration counts of the benchmark batch under two different random inits
(is the length of a run a property of the spectrum or of the starting point?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from bayes_drt_b200 import capi, synth
from oracle import model as omod
B = 12500
freq, Z, _ = synth.make_spectra(B, seed=1)
_, bf = synth.bench_grid()
tau = 1 / (2 * np.pi * bf.numpy()); eps = omod.default_epsilon(tau)
A_re, A_im = capi.build_A(freq, tau, eps)
c = omod.MODE_CONSTANTS['optimize']
bft = torch.tensor(1/(2*np.pi*tau))
L = torch.stack([c[f'l{o}'] * capi.build_L(bft, torch.tensor(tau), eps, o) for o in (0,1,2)])
Zc = Z.cuda()
zs = (Zc.abs().std(dim=1, unbiased=False) / np.sqrt(70/81))
Zst = torch.cat(((Zc / zs[:, None]).real, (Zc / zs[:, None]).imag), dim=1).contiguous()
prob = capi.SeriesProblem(torch.cat((A_re, A_im)), Zst, freq, L)
out = {}
for s in (0, 5):
    g = torch.Generator().manual_seed(s)
    u0 = (torch.rand(B, prob.D, generator=g, dtype=torch.float64) * 4 - 2).cuda()
    r = prob.map_lbfgs(u0, max_iter=50000)
    for k in ('iters', 'n_eval', 'lp', 'status'):
        out[f'{k}_{s}'] = r[k].cpu().numpy()
np.savez_compressed('gpurun_out/iters.npz', **out)
a, b = out['iters_0'].astype(float), out['iters_5'].astype(float)
print('corr(iters init0, iters init5) =', np.corrcoef(a, b)[0, 1], ' rank corr', np.corrcoef(np.argsort(np.argsort(a)), np.argsort(np.argsort(b)))[0, 1])
