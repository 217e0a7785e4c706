import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayes_drt_b200 import capi, synth
from oracle import model as omod
B = 1184
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 1184 * 200
freq, Z, _ = synth.make_spectra(B, seed=1)
_, bf = synth.bench_grid()
tau = 1 / (2 * np.pi * bf.numpy()); eps = omod.default_epsilon(tau)
A_re, A_im = capi.build_A(freq, tau, eps)
c = omod.MODE_CONSTANTS['optimize']
bft = torch.tensor(1/(2*np.pi*tau))
L = torch.stack([c[f'l{o}'] * capi.build_L(bft, torch.tensor(tau), eps, o) for o in (0,1,2)])
Zc = Z.cuda(); zs = (Zc.abs().std(dim=1, unbiased=False) / np.sqrt(70/81)); Zs = Zc / zs[:, None]
Zst = torch.cat((Zs.real, Zs.imag), dim=1).contiguous()
for nonneg, outl in ((False, False), (True, True)):
    prob = capi.SeriesProblem(torch.cat((A_re, A_im)), Zst, freq, L, nonneg=nonneg, outliers=outl)
    g = torch.Generator().manual_seed(0)
    u = (torch.rand(ncol, prob.D, generator=g, dtype=torch.float64) - 0.5).cuda()
    for rep in range(3):
        torch.cuda.synchronize(); t = time.time()
        lp, grad = prob.logpost_grad(u)
        torch.cuda.synchronize(); dt = time.time() - t
    print(f'model pos={nonneg} out={outl} D={prob.D}: {ncol} grads in {dt*1e3:.2f} ms -> {ncol/dt:.3e} grads/s, {ncol*84000/dt/1e12:.2f} TFLOP/s (banded), round {dt/(ncol/1184)*1e6:.2f} us')
