"""Small drivers for ncu captures of the kernels off the main solvers: build_A_kernel (per-spectrum grids, config 5
shape) and newton_kernel (polish of 296 L-BFGS end points of the benchmark shape).  python scripts/gpu_time_misc.py A|newton"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayes_drt_b200 import Inverter, capi, synth
what = sys.argv[1] if len(sys.argv) > 1 else 'A'
if what == 'A':
    Gm = 2048
    fg = 10.0 ** (6.0 - torch.rand(Gm, 1, dtype=torch.float64) - torch.arange(81, dtype=torch.float64)[None, :] / 10.0)
    taum = torch.as_tensor(1.0 / (2 * np.pi * np.logspace(6, -2, 81)))
    eps = 1.0 / float(np.mean(np.diff(np.log(taum.numpy()))))
    fgd = fg.cuda()
    capi.build_A(fgd[:8], taum, eps)
    torch.cuda.synchronize(); t = time.time()
    capi.build_A(fgd, taum, eps)
    torch.cuda.synchronize(); dt = time.time() - t
    print(f'build_A: {2 * Gm / dt:.0f} matrices/s ({Gm} grids, 81 x 81, both parts)')
else:
    B = 296
    freq, Z, _ = synth.make_spectra(B, seed=1)
    _, bf = synth.bench_grid()
    inv = Inverter(basis_freq=bf.numpy())
    prob, u0 = inv.prepare(freq, Z, mode='optimize')
    r = prob.map_lbfgs(u0, max_iter=50000)
    prob.map_newton(r['u'][:8], max_iter=2) if False else None
    torch.cuda.synchronize(); t = time.time()
    p = prob.map_newton(r['u'])
    torch.cuda.synchronize(); dt = time.time() - t
    print(f'newton: {B / dt:.1f} polishes/s, iterations mean {p["iters"].float().mean().item():.1f}, '
          f'gradients mean {p["n_eval"].float().mean().item():.0f}, converged {(p["gnorm"] < 2e-9).float().mean().item():.3f}')
