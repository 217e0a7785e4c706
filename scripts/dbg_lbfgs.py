import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from helpers import gpu_problem, load_spectrum, oracle_batch
from oracle import lbfgs as olb, model as omod
np.seterr(all='ignore')
freq, Z = load_spectrum('ZARC_uniform_0.25')
ds = oracle_batch(freq, [Z], mode='optimize')
prob = gpu_problem(ds)
rng = np.random.RandomState(0)
u0 = rng.uniform(-2, 2, (2, prob.D))[:1]
pts = []
def f(u):
    lp, g = omod.logpost(u, ds[0])
    ok = np.isfinite(lp) and np.all(np.isfinite(g))
    pts.append((u.copy(), lp, g.copy(), ok))
    return (-lp, -g) if ok else None
for n_it in (1, 2, 3):
    pts.clear()
    o = olb.minimize(f, u0[0], max_iter=n_it)
    r = prob.map_lbfgs(torch.tensor(u0), max_iter=n_it)
    print('n_it', n_it, 'oracle f', o['f'], 'nev', o['n_eval'], 'gpu lp', r['lp'].item(), 'nev', r['n_eval'].item(), 'status', r['status'].item(),
          'udiff', np.max(np.abs(r['u'][0].cpu().numpy() - o['x'])))
    U = torch.tensor(np.stack([p[0] for p in pts]))
    lp, g = prob.logpost_grad(U, spec=np.zeros(len(pts), dtype=np.int32))
    for i, p in enumerate(pts):
        gg = g[i].cpu().numpy()
        okg = np.isfinite(lp[i].item()) and np.all(np.isfinite(gg))
        print('  eval', i, 'oracle lp %.6e ok %d | gpu lp %.6e ok %d | step %.3e' % (p[1], p[3], lp[i].item(), okg, np.linalg.norm(p[0]-u0[0])))
