"""Per-section clocks of ridge_kernel from a profiling build: BDRT_LIB=scratch_libs/libbdrt_clk.so python scripts/gpu_ridge_clocks.py"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayes_drt_b200 import Inverter, _lib, synth
B = 4096
freq, Z, _ = synth.make_spectra(B, seed=1)
_, bf = synth.bench_grid()
ctx = _lib.context()
names = {1: 'assemble', 2: 'cholesky', 3: 'solves', 4: 'multipliers/exchange', 5: 'lambda', 6: 'penalty', 7: 'gram+q', 8: 'stop'}
for kw in (dict(stop_rule='nan'), dict(preset='Huang')):
    inv = Inverter(basis_freq=bf.numpy())
    inv.ridge_fit(freq, Z[:64], **kw)
    out = (C.c_ulonglong * 16)()
    ctx.lib.bdrt_debug_phase_clocks(ctx._h, out)
    torch.cuda.synchronize(); t = time.time()
    inv.ridge_fit(freq, Z, **kw)
    torch.cuda.synchronize(); dt = time.time() - t
    ctx.lib.bdrt_debug_phase_clocks(ctx._h, out)
    npiv = max(out[0], 1)
    print(kw, f'{B/dt:.0f} fits/s, pivots/fit {npiv/B:.1f}; clocks per pivot: ' +
          ' | '.join(f'{names[i]} {out[i]/npiv:.0f}' for i in range(1, 9)), f'total {sum(out[1:9])/npiv:.0f}')
