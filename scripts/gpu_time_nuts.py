import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from bayes_drt_b200 import capi, synth
from oracle import model as omod
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 2
warmup = int(sys.argv[3]) if len(sys.argv) > 3 else 200
samples = int(sys.argv[4]) if len(sys.argv) > 4 else 200
freq, Z, _ = synth.make_spectra(B, seed=1)
_, bf = synth.bench_grid()
tau = 1 / (2 * np.pi * bf.numpy()); eps = omod.default_epsilon(tau)
A_re, A_im = capi.build_A(freq, tau, eps)
c = omod.MODE_CONSTANTS['sample']
bft = torch.tensor(1/(2*np.pi*tau))
L = torch.stack([c[f'l{o}'] * capi.build_L(bft, torch.tensor(tau), eps, o) for o in (0,1,2)])
Zc = Z.cuda(); zs = (Zc.abs().std(dim=1, unbiased=False) / np.sqrt(70/81)); Zs = Zc / zs[:, None]
Zst = torch.cat((Zs.real, Zs.imag), dim=1).contiguous()
prob = capi.SeriesProblem(torch.cat((A_re, A_im)), Zst, freq, L, ups_alpha=c['ups_alpha'], ups_beta=c['ups_beta'])
g = torch.Generator().manual_seed(0)
u0 = (torch.rand(B, chains, prob.D, generator=g, dtype=torch.float64) * 4 - 2).cuda()
torch.cuda.synchronize(); t = time.time()
r = prob.nuts(u0, chains=chains, warmup=warmup, samples=samples)
torch.cuda.synchronize(); dt = time.time() - t
ng = r['n_leapfrog'].sum().item()
print(f'B={B} chains={chains} {warmup}+{samples}: {dt:.2f}s  inversions/s {B/dt:.2f}  grads {ng:.3e} grads/s {ng/dt:.3e} TFLOP/s {ng*84000/dt/1e12:.2f}')
print('stepsize median', r['stepsize'].median().item(), 'accept mean', r['accept'].mean().item(), 'div', r['n_divergent'].sum().item(), 'maxdepth frac', r['n_maxdepth'].float().mean().item()/max(samples,1), 'leap/iter', ng/(B*chains*(warmup+samples)))
print('finite', torch.isfinite(r['draws']).all().item())
