"""Capture the Stan input the reference's OWN Inverter.fit builds (bayes_drt/inversion.py imported unmodified from
/root/reference: _prep_matrices, _scale_Z, _get_stan_model, _prep_stan_data, the outlier N override, init_from_ridge)
and store it as golden vectors for the oracle's data preparation -> tests/golden/stan_data.npz.

The reference's natives are absent in the build container; inside this script only they are replaced by stubs:
load_pickle returns a fake StanModel whose optimizing()/sampling() record the model file name, the data dict and the
init and then abort the fit; cvxopt.solvers.qp is an exact QP solver (only reached by init_from_ridge); matplotlib is an
empty module.  Matrices are stored as random projections (M v, u M) -- the matrices themselves are pinned entry by
entry in matrices.npz.
Run:  PYTHONPATH=/root/reference python scripts/make_golden_stan_data.py
"""
import os
import sys
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ridge as oridge  # noqa: E402  (qp_bound only, as the stand-in for cvxopt)

cv = types.ModuleType('cvxopt')
cv.matrix = lambda a: np.array(a, dtype=np.float64)
cv.solvers = types.SimpleNamespace(options={})


def _qp(P, q, G, h):
    P, q, G, h = (np.asarray(a, dtype=np.float64) for a in (P, q, G, h))
    x, y, F, it = oridge.qp_bound(P.T, q.ravel(), -h.ravel(), strict=False)
    return {'x': x.copy(), 'primal objective': 0.0, 'status': 'optimal'}


cv.solvers.qp = _qp
sys.modules['cvxopt'] = cv
captured = {}


class Abort(Exception):
    pass


class FakeModel:
    def __init__(self, name):
        self.name = name

    def optimizing(self, dat, iter=None, seed=None, init=None):
        captured.clear()
        captured.update(model=self.name, dat=dat, iter=iter, seed=seed, init=init)
        raise Abort

    def sampling(self, dat, **kw):
        captured.clear()
        captured.update(model=self.name, dat=dat, **kw)
        raise Abort


sm = types.ModuleType('bayes_drt.stan_models')
sm.save_pickle = lambda o, f: None
sm.load_pickle = lambda f: FakeModel(os.path.basename(f))
sys.modules['bayes_drt.stan_models'] = sm
mpl = types.ModuleType('matplotlib')
mpl.pyplot = types.ModuleType('matplotlib.pyplot')
sys.modules['matplotlib'] = mpl
sys.modules['matplotlib.pyplot'] = mpl.pyplot
warnings.simplefilter('ignore')
from bayes_drt.inversion import Inverter  # noqa: E402

g = np.load(os.path.join(ROOT, 'tests', 'golden', 'spectra.npz'))
freq, Z = g['ZARC_uniform_0.25/freq'], g['ZARC_uniform_0.25/Z']
bf = np.logspace(6, -2, 81)
TP = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'transmissive', 'dist_type': 'parallel', 'basis_freq': bf}
BP = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'blocking', 'dist_type': 'parallel', 'basis_freq': bf}
DRT = {'kernel': 'DRT', 'basis_freq': bf}
CASES = {
    'series_opt': (dict(), dict(mode='optimize')),
    'series_sample': (dict(), dict(mode='sample')),
    'series_pos_opt': (dict(), dict(mode='optimize', nonneg=True)),
    'series_out_opt': (dict(), dict(mode='optimize', outliers=True)),
    'series_pos_out_sample': (dict(), dict(mode='sample', nonneg=True, outliers=True, outlier_lambda=5, sigma_min=0.001,
                                           inductance_scale=2)),
    'series_basis_eq_freq': (dict(basis_freq=freq), dict(mode='optimize')),
    'series_noscale': (dict(), dict(mode='optimize', scale_Z=False)),
    'series_ridge_init': (dict(), dict(mode='optimize', init_from_ridge=True)),
    'series_pos_out_ridge_init': (dict(), dict(mode='sample', nonneg=True, outliers=True, init_from_ridge=True)),
    'sp_opt': (dict(distributions={'DRT': dict(DRT), 'TP-DDT': dict(TP, x_scale=0.8)}), dict(mode='optimize', nonneg=True)),
    'sp_sample': (dict(distributions={'DRT': dict(DRT), 'TP-DDT': dict(TP, x_scale=0.8)}), dict(mode='sample', nonneg=True)),
    'parallel_tp_opt': (dict(distributions={'TP-DDT': dict(TP)}), dict(mode='optimize')),
    'parallel_bp_sample': (dict(distributions={'BP-DDT': dict(BP)}), dict(mode='sample')),
    's2p_opt': (dict(distributions={'DRT': dict(DRT), 'TP-DDT': dict(TP, x_scale=0.8), 'BP-DDT': dict(BP)}),
                dict(mode='optimize', nonneg=True)),
    's2p_sample': (dict(distributions={'DRT': dict(DRT), 'TP-DDT': dict(TP, x_scale=0.8), 'BP-DDT': dict(BP)}),
                   dict(mode='sample', nonneg=True)),
}
# outliers='auto' (inversion.py:1171-1187) on a spectrum with two gross outliers (made by make_golden_ridge_reference.py)
Zo = np.load(os.path.join(ROOT, 'tests', 'golden', 'ridge_reference.npz'))['outliers/Z']
CASES['auto_contaminated'] = (dict(), dict(mode='optimize', outliers='auto'))
CASES['auto_contaminated_ridge_init'] = (dict(), dict(mode='optimize', outliers='auto', init_from_ridge=True))
CASES['auto_clean_ridge_init'] = (dict(), dict(mode='sample', outliers='auto', init_from_ridge=True))
# a third, marginal outlier: flagged at the initialisation's threshold 3 (inversion.py:1668-1675) but not at the model
# choice's threshold 4 (:1171-1187)
Zm = Zo.copy()
Zm[70] = Z[70] + 0.04
CASES['auto_marginal_ridge_init'] = (dict(), dict(mode='optimize', outliers='auto', init_from_ridge=True))
CASES['true_marginal_ridge_init'] = (dict(), dict(mode='optimize', outliers=True, init_from_ridge=True))
rng = np.random.RandomState(5)
out = {'freq': freq, 'Z': Z, 'Z_contaminated': Zo, 'Z_marginal': Zm}
for case, (ikw, fkw) in CASES.items():
    inv = Inverter(**ikw)
    try:
        inv.fit(freq, Zo if 'contaminated' in case else (Zm if 'marginal' in case else Z), **fkw)
    except Abort:
        pass
    dat = captured['dat']
    out[f'{case}/model'] = np.array(captured['model'])
    out[f'{case}/Z_scale'] = np.float64(inv._Z_scale)
    for k, v in dat.items():
        if k.endswith('_tilde'):
            continue
        v = np.asarray(v, dtype=np.float64)
        if v.ndim == 2:  # projections with fixed random vectors (seeded per shape)
            r = np.random.RandomState(v.shape[0] * 1000 + v.shape[1])
            out[f'{case}/{k}@v'] = v @ r.standard_normal(v.shape[1])
            out[f'{case}/u@{k}'] = r.standard_normal(v.shape[0]) @ v
        else:
            out[f'{case}/{k}'] = v
    init = captured.get('init')
    if callable(init):
        d = init()
        for k, v in d.items():
            out[f'{case}/init/{k}'] = np.asarray(v, dtype=np.float64)
    else:
        out[f'{case}/init'] = np.array(str(init))
    for k in ('iter', 'seed', 'warmup', 'chains'):
        if captured.get(k) is not None:
            out[f'{case}/{k}'] = np.int64(captured[k])
    if 'control' in captured:
        out[f'{case}/control'] = np.array([captured['control']['adapt_delta'], captured['control']['adapt_t0']])
    print(case, captured['model'], 'Z_scale %.6f' % inv._Z_scale, sorted(k for k in dat if not k.endswith('_tilde'))[:6], '...')
dst = os.path.join(ROOT, 'tests', 'golden', 'stan_data.npz')
np.savez_compressed(dst, **out)
print('wrote', dst, os.path.getsize(dst), 'bytes')
