import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np, torch
from test_gpu_inverter import _zarc_batch
from bayes_drt_b200 import Inverter
freq, Z = _zarc_batch([0.0, 0.13, 0.37, 0.5, 0.82])
for polish in (False, True):
    inv = Inverter(); inv.fit(freq, Z, mode='optimize', polish=polish)
    r = inv._opt_result
    print('batch polish', polish, {k: (v.cpu().numpy() if torch.is_tensor(v) and v.numel() <= 8 else None) for k, v in r.items()})
    for b in range(5):
        one = Inverter(); one.fit(freq[b], Z[b], mode='optimize', polish=polish, spectrum_offset=b)
        r1 = one._opt_result
        c1 = one.distribution_fits['DRT']['coef']; cb = inv.distribution_fits['DRT']['coef'][b].cpu().numpy()
        print(' single', b, {k: (v.cpu().numpy() if torch.is_tensor(v) and v.numel() <= 8 else None) for k, v in r1.items()},
              'dcoef', np.abs(c1 - cb).max() / np.abs(c1).max())
# order dependence: reversed batch
inv = Inverter(); inv.fit(freq[::-1].copy(), Z[::-1].copy(), mode='optimize', polish=True)
print('reversed', inv._opt_result['gnorm'].cpu().numpy(), inv._opt_result['iters'].cpu().numpy(), inv._opt_result['lp'].cpu().numpy())
