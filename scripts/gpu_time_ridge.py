"""Throughput of the hyper-lambda ridge kernel on the benchmark shape, with the factorisations counted on device."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayes_drt_b200 import Inverter, capi, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
freq, Z, _ = synth.make_spectra(B, seed=1)
_, bf = synth.bench_grid()
dfma, dmma = capi.peak_fp64()
Inverter(basis_freq=bf.numpy()).ridge_fit(freq, Z)  # first full-size launch: workspace growth, module load
for kw in (dict(), dict(stop_rule='nan'), dict(preset='Huang'), dict(penalty='integral', lambda_0=1, hl_beta=5, weights='modulus')):
    inv = Inverter(basis_freq=bf.numpy())
    inv.ridge_fit(freq, Z[:64], **kw)
    torch.cuda.synchronize(); t = time.time()
    inv.ridge_fit(freq, Z, **kw)
    torch.cuda.synchronize(); dt = time.time() - t
    n, Nf = 102, 70
    it = inv._ridge_iters.float().mean().item()
    # the library counts the Cholesky factorisations; flop model: Gram matrix 2 (2 Nf) n^2 when the weights are
    # per spectrum, per hyper-iteration the penalty assembly 2 n^2 (+ 2 n^2 lambda update), per factorisation n^3 / 3 + 4 n^2
    fac = inv._ridge_factorisations.float().mean().item()
    flop = (2.0 * 2 * Nf * n * n if kw.get('weights') or kw.get('preset') else 0.0) + it * 4.0 * n * n + fac * (n ** 3 / 3.0 + 4.0 * n * n)
    print(f'{kw}: {B} fits in {dt*1e3:.1f} ms -> {B/dt:.0f} fits/s; hyper-iterations {it:.2f}, factorisations {fac:.2f}, '
          f'{B*flop/dt/1e12:.2f} TFLOP/s = {100*B*flop/dt/1e12/dmma:.1f} % of the DMMA peak ({dmma:.1f}), converged {inv._ridge_converged.float().mean().item():.3f}')
