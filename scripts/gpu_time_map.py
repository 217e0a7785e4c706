import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from bayes_drt_b200 import capi, synth
from oracle import model as omod
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
freq, Z, _ = synth.make_spectra(B, seed=1)
_, bf = synth.bench_grid()
tau = 1 / (2 * np.pi * bf.numpy()); eps = omod.default_epsilon(tau)
A_re, A_im = capi.build_A(freq, tau, eps)
c = omod.MODE_CONSTANTS['optimize']
bft = torch.tensor(1/(2*np.pi*tau))
L = torch.stack([c[f'l{o}'] * capi.build_L(bft, torch.tensor(tau), eps, o) for o in (0,1,2)])
Zc = Z.cuda()
zs = (Zc.abs().std(dim=1, unbiased=False) / np.sqrt(70/81))
Zs = Zc / zs[:, None]
Zst = torch.cat((Zs.real, Zs.imag), dim=1).contiguous()
prob = capi.SeriesProblem(torch.cat((A_re, A_im)), Zst, freq, L)
g = torch.Generator().manual_seed(0)
u0 = (torch.rand(B, prob.D, generator=g, dtype=torch.float64) * 4 - 2).cuda()
for rep in range(2):
    torch.cuda.synchronize(); t = time.time()
    r = prob.map_lbfgs(u0, max_iter=max_iter)
    torch.cuda.synchronize(); dt = time.time() - t
    nev = r['n_eval'].sum().item()
    print(f'B={B} max_iter={max_iter} time {dt:.3f}s  spectra/s {B/dt:.1f}  grads {nev}  grads/s {nev/dt:.3e}  TFLOP/s(banded 84k) {nev*84000/dt/1e12:.2f}')
st = r['status'].cpu().numpy()
print('status counts', {int(k): int((st==k).sum()) for k in np.unique(st)}, 'iters mean', r['iters'].float().mean().item(), 'max', r['iters'].max().item(), 'nev mean', r['n_eval'].float().mean().item())
print('lp mean', r['lp'].mean().item())
