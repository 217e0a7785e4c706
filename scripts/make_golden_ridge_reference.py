"""Run the reference's OWN Inverter.ridge_fit (bayes_drt/inversion.py, imported unmodified from /root/reference) and
store its results as golden vectors for the oracle's ridge restatement -> tests/golden/ridge_reference.npz.

The reference's third-party natives are absent in the build container (cvxopt, pystan, matplotlib).  They are replaced
by stubs that exist only inside this script:
  * cvxopt.matrix -> numpy array, cvxopt.solvers.qp(P, q, G, h) -> an EXACT solver of the same bound-constrained QP
    (G = -I, h = -lower bounds; oracle.ridge.qp_bound).  Everything else -- scaling, weights, matrices, the hyper-lambda
    loop, its stop test, Re-Im cross-validation, presets, the least-squares offsets of one-part fits, rescaling -- is
    the reference's code, line by line.  So the fixtures pin the oracle's restatement of ridge_fit up to the QP solver
    (cvxopt's interior-point iterates differ from the exact solution by its tolerances 1e-7 / 1e-6).
  * bayes_drt.stan_models -> pickle helpers only (the real module compiles Stan programs at import).
  * matplotlib -> empty module (plotting is never called).
Run:  PYTHONPATH=/root/reference python scripts/make_golden_ridge_reference.py
"""
import os
import sys
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ridge as oridge  # noqa: E402  (only qp_bound is used, as the stand-in for cvxopt)

# ---- stubs ---------------------------------------------------------------------------------------------------------
cv = types.ModuleType('cvxopt')
cv.matrix = lambda a: np.array(a, dtype=np.float64)
cv.solvers = types.SimpleNamespace(options={})


def _qp(P, q, G, h):
    P, q, G, h = (np.asarray(a, dtype=np.float64) for a in (P, q, G, h))
    assert np.array_equal(G, -np.eye(len(q)))  # x >= -h
    x, y, F, it = oridge.qp_bound(P.T, q.ravel(), -h.ravel(), strict=False)
    return {'x': x.copy(), 'primal objective': float(0.5 * x @ P.T @ x + q.ravel() @ x), 'status': 'optimal'}


cv.solvers.qp = _qp
sys.modules['cvxopt'] = cv
sm = types.ModuleType('bayes_drt.stan_models')
sm.save_pickle = lambda obj, file: None
sm.load_pickle = lambda file: None
sys.modules['bayes_drt.stan_models'] = sm
mpl = types.ModuleType('matplotlib')
mpl.pyplot = types.ModuleType('matplotlib.pyplot')
sys.modules['matplotlib'] = mpl
sys.modules['matplotlib.pyplot'] = mpl.pyplot

warnings.simplefilter('ignore')
from bayes_drt.inversion import Inverter  # noqa: E402

g = np.load(os.path.join(ROOT, 'tests', 'golden', 'spectra.npz'))
CASES = {
    'default': dict(),
    'huang': dict(preset='Huang'),
    'init_from_ridge': dict(penalty='integral', hyper_lambda=True, lambda_0=1, hl_beta=5, weights='modulus'),
    'free_sign': dict(nonneg=False),
    'mixed_orders': dict(reg_ord=[0.2, 0.3, 0.5], L1_penalty=0.01),
    'real_part': dict(part='real'),
    'imag_part': dict(part='imag', weights='modulus'),
    'fbeta': dict(hl_fbeta=0.1),
    'cholesky': dict(penalty='cholesky'),
    'cv': dict(lambda_0='cv', cv_lambdas=np.logspace(-6, 0, 7)),
    'ciucci': dict(preset='Ciucci', cv_lambdas=np.logspace(-6, 0, 7)),
}
out = {}
for name in ('ZARC_uniform_0.25', '2ZARC_uniform_0.25'):
    freq, Z = g[name + '/freq'], g[name + '/Z']
    for case, kw in CASES.items():
        inv = Inverter()
        inv.ridge_fit(freq, Z, **kw)
        key = f'{name}/{case}'
        out[key + '/coef'] = np.asarray(inv.distribution_fits['DRT']['coef'], dtype=np.float64)
        out[key + '/R_inf'] = np.float64(inv.R_inf)
        out[key + '/inductance'] = np.float64(inv.inductance)
        if hasattr(inv, 'cv_result') and 'cv' in str(kw.get('lambda_0', '')) + str(kw.get('preset', '')).replace('Ciucci', 'cv'):
            out[key + '/cv_result'] = inv.cv_result.values.astype(np.float64)
        print(key, 'R_inf %.6f' % inv.R_inf, 'sum coef %.6f' % out[key + '/coef'].sum())
# check_outliers without a Stan fit (inversion.py:3313-3367): ridge fit with preset 'Huang' + inter-quartile rule
freq, Z = g['ZARC_uniform_0.25/freq'], g['ZARC_uniform_0.25/Z']
Zo = Z.copy()
Zo[30] += 0.2 + 0.2j
Zo[55] -= 0.1j
Zo[70] += 0.02
out['outliers/Z'] = Zo
for thr in (0.5, 1.5, 4):
    inv = Inverter()
    out[f'outliers/idx_t{thr}'] = np.asarray(inv.check_outliers(freq, Zo, threshold=thr, use_existing_fit=False)).ravel()
    print('outliers at threshold', thr, out[f'outliers/idx_t{thr}'])
dst = os.path.join(ROOT, 'tests', 'golden', 'ridge_reference.npz')
np.savez_compressed(dst, **out)
print('wrote', dst, os.path.getsize(dst), 'bytes')
