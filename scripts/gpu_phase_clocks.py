"""Per-phase clock breakdown of engine_eval inside logpost_kernel / lbfgs_kernel / nuts_kernel, from a profiling build
(bash bayes_drt_b200/csrc/build.sh -DBDRT_PHASE_CLOCKS -> scratch_libs/libbdrt_clk.so):

    BDRT_LIB=$PWD/scratch_libs/libbdrt_clk.so python scripts/gpu_phase_clocks.py

Every warp adds the clock64() difference of each phase of each evaluation to a global counter; printed per evaluation."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bayes_drt_b200 import _lib, capi, synth
from oracle import model as omod

B = 4736
freq, Z, _ = synth.make_spectra(B, seed=1)
_, bf = synth.bench_grid()
tau = 1 / (2 * np.pi * bf.numpy()); eps = omod.default_epsilon(tau)
A_re, A_im = capi.build_A(freq, tau, eps)
bft = torch.tensor(1 / (2 * np.pi * tau))
Zc = Z.cuda(); zs = (Zc.abs().std(dim=1, unbiased=False) / np.sqrt(70 / 81)); Zs = Zc / zs[:, None]
Zst = torch.cat((Zs.real, Zs.imag), dim=1).contiguous()
ctx = _lib.context()


def clocks(label, n_warp_slots, dt):
    out = (C.c_ulonglong * 16)()
    rc = ctx.lib.bdrt_debug_phase_clocks(ctx._h, out)
    assert rc == 0, rc
    n = out[0]
    ph = [out[i] / max(n, 1) for i in range(1, 6)]
    tot = sum(ph)
    print(f'{label}: {n} evals in {dt*1e3:.1f} ms; clocks/eval in engine {tot:.0f} = ' +
          ' | '.join(f'p{i+1} {p:.0f} ({100*p/tot:.0f}%)' for i, p in enumerate(ph)) +
          f' ; wall clocks per eval per warp {dt*1.965e9*n_warp_slots/max(n,1):.0f}')
    if any(out[i] for i in range(6, 12)):
        names = {6: 'linesearch+step', 7: 'engine', 8: 'grad post-pass', 9: 'shift+sweepA', 10: 'recursion', 11: 'sweepB'}
        print('   solver clocks/eval: ' + ' | '.join(f'{names[i]} {out[i]/max(n,1):.0f}' for i in range(6, 12)))


for mode in ('optimize', 'sample'):
    c = omod.MODE_CONSTANTS[mode]
    L = torch.stack([c[f'l{o}'] * capi.build_L(bft, torch.tensor(tau), eps, o) for o in (0, 1, 2)])
    prob = capi.SeriesProblem(torch.cat((A_re, A_im)), Zst, freq, L, ups_alpha=c['ups_alpha'], ups_beta=c['ups_beta'])
    g = torch.Generator().manual_seed(0)
    if mode == 'optimize':
        u = (torch.rand(1184 * 200, prob.D, generator=g, dtype=torch.float64) - 0.5).cuda()
        prob.logpost_grad(u); clocks('warm', 1, 1)
        torch.cuda.synchronize(); t = time.time(); prob.logpost_grad(u); torch.cuda.synchronize()
        clocks('logpost_kernel', 148 * 16, time.time() - t)
        u0 = (torch.rand(B, prob.D, generator=g, dtype=torch.float64) * 4 - 2).cuda()
        torch.cuda.synchronize(); t = time.time(); r = prob.map_lbfgs(u0, max_iter=2000); torch.cuda.synchronize()
        clocks('lbfgs_kernel (4736 x 2000 its)', 148 * 16, time.time() - t)
    else:
        prob = capi.SeriesProblem(torch.cat((A_re, A_im)), Zst[:1184], freq, L, ups_alpha=c['ups_alpha'], ups_beta=c['ups_beta'])
        u0 = (torch.rand(1184, 2, prob.D, generator=g, dtype=torch.float64) * 4 - 2).cuda()
        torch.cuda.synchronize(); t = time.time(); r = prob.nuts(u0, chains=2, warmup=60, samples=20); torch.cuda.synchronize()
        clocks('nuts_kernel (1184 x 2 x (60+20))', 148 * 16, time.time() - t)
