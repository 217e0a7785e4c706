"""Golden vectors of the reference's result extraction and post-fit queries (inversion.py:1222-1289, :2494-2566,
:2669-3160, :3162-3311): Inverter.fit is run from the unmodified reference with a fake StanModel that returns a
SYNTHETIC Stan result (seeded random numbers of plausible size), so everything recorded afterwards -- coefficient
rescaling, R_inf / inductance / error-model extraction, predict_distribution / predict_Z / predict_Rp / predict_sigma /
coef_percentile / score / check_outliers -- is the reference's own arithmetic on known inputs.
-> tests/golden/postfit.npz.   Run:  PYTHONPATH=/root/reference python scripts/make_golden_postfit.py
(stubs as in make_golden_stan_data.py: cvxopt, pystan via bayes_drt.stan_models, matplotlib are absent here)."""
import os
import sys
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cv = types.ModuleType('cvxopt')
cv.matrix = lambda a: np.array(a, dtype=np.float64)
cv.solvers = types.SimpleNamespace(options={}, qp=None)
sys.modules['cvxopt'] = cv
STATE = {}


def synthetic(dat, n_draws, seed):
    """A Stan result for the data dict `dat`: every quantity the reference reads, [n_draws, ...] if n_draws else [...]."""
    r = np.random.RandomState(seed)
    lead = (n_draws,) if n_draws else ()
    N2 = len(dat['Z'])
    out = {'Rinf': 2.0 + 0.1 * r.rand(*lead), 'induc': 1e-7 * r.rand(*lead), 'sigma_res': 0.01 + 0.01 * r.rand(*lead),
           'alpha_prop': 0.02 * r.rand(*lead), 'alpha_re': 0.01 * r.rand(*lead), 'alpha_im': 0.015 * r.rand(*lead),
           'sigma_tot': 0.01 + 0.02 * r.rand(*lead, N2)}
    if 'K' in dat:
        out['x'] = 0.3 * r.rand(*lead, dat['K']) if 'Parallel' in STATE['model'] else 0.2 * r.randn(*lead, dat['K'])
    else:
        out['xs'] = 0.2 * r.rand(*lead, dat['Ks'])
        for k in ('Kp', 'Kp1', 'Kp2'):
            if k in dat:
                out['x' + k[1:]] = 0.3 * r.rand(*lead, dat[k])
    if 'sigma_out_lambda' in dat:
        out['sigma_out'] = 0.05 * r.rand(*lead, dat['N'])
    if n_draws:
        Nf = N2 // 2
        out['Z_hat'] = r.randn(n_draws, N2)  # present in Stan's generated quantities; not used below
    return out


class FitResult(dict):
    """StanFit4Model raises ValueError for an unknown parameter name (predict_sigma relies on it, inversion.py:3112-3115)"""
    def __getitem__(self, k):
        if k not in self:
            raise ValueError(f'No parameter {k}')
        return dict.__getitem__(self, k)


class FakeModel:
    def __init__(self, name):
        self.name = name

    def optimizing(self, dat, iter=None, seed=None, init=None):
        STATE.update(model=self.name, dat=dat)
        STATE['result'] = synthetic(dat, 0, 7)
        return STATE['result']

    def sampling(self, dat, **kw):
        STATE.update(model=self.name, dat=dat)
        STATE['result'] = FitResult(synthetic(dat, 40, 8))
        return STATE['result']


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
    'series_out_opt': (dict(), dict(mode='optimize', outliers=True)),
    'series_sample': (dict(), dict(mode='sample')),
    'series_out_sample': (dict(), dict(mode='sample', outliers=True)),
    'parallel_opt': (dict(distributions={'TP-DDT': dict(TP)}), dict(mode='optimize')),
    'parallel_sample': (dict(distributions={'TP-DDT': dict(TP)}), dict(mode='sample')),
    'sp_opt': (dict(distributions={'DRT': dict(DRT), 'TP-DDT': dict(TP, x_scale=0.8)}), dict(mode='optimize', nonneg=True)),
    'sp_sample': (dict(distributions={'DRT': dict(DRT), 'TP-DDT': dict(TP, x_scale=0.8)}), dict(mode='sample', nonneg=True)),
    's2p_opt': (dict(distributions={'DRT': dict(DRT), 'TP-DDT': dict(TP, x_scale=0.8), 'BP-DDT': dict(BP)}),
                dict(mode='optimize', nonneg=True)),
}
f_pred = np.sort(10.0 ** np.random.RandomState(3).uniform(-1.5, 5.5, 81))  # 81 points (the reference compares f_pred with
# the previous grid element-wise), ascending and irregular (its Toeplitz shortcut mis-slices other log-uniform grids)
eval_tau = np.logspace(-7, 2, 37)
out = {'freq': freq, 'Z': Z, 'f_pred': f_pred, 'eval_tau': eval_tau}
for case, (ikw, fkw) in CASES.items():
    inv = Inverter(**ikw)
    inv.fit(freq, Z, check_outliers=False, **fkw)
    p = case + '/'
    out[p + 'model'] = np.array(STATE['model'])
    out[p + 'Z_scale'] = np.float64(inv._Z_scale)
    for k, v in STATE['result'].items():
        out[p + 'stan/' + k] = np.asarray(v, dtype=np.float64)
    for name, d in inv.distribution_fits.items():
        out[p + 'coef/' + name] = np.asarray(d['coef'], dtype=np.float64)
        out[p + 'gamma/' + name] = inv.predict_distribution(name, eval_tau=eval_tau)
        out[p + 'tau/' + name] = np.asarray(inv.distributions[name]['tau'], dtype=np.float64)
        out[p + 'epsilon/' + name] = np.float64(inv.distributions[name]['epsilon'])
        if fkw['mode'] == 'sample':
            out[p + 'coef_p2.5/' + name] = inv.coef_percentile(name, 2.5)
            out[p + 'gamma_p90/' + name] = inv.predict_distribution(name, eval_tau=eval_tau, percentile=90)
    out[p + 'R_inf'], out[p + 'inductance'] = np.float64(inv.R_inf), np.float64(inv.inductance)
    for k, v in inv.error_fit.items():
        out[p + 'error_fit/' + k] = np.asarray(v, dtype=np.float64)
    out[p + 'Z_pred_train'] = inv.predict_Z(freq)
    out[p + 'Z_pred'] = inv.predict_Z(f_pred)
    out[p + 'Z_pred_no_offsets'] = inv.predict_Z(f_pred, include_offsets=False)
    # predict_Rp: closed form for a single DRT; for the other models the reference predicts Z at two frequencies and
    # compares that grid element-wise with the 81 training frequencies, which numpy >= 1.25 refuses -- not recorded
    single_drt = len(inv.distributions) == 1 and list(inv.distributions.values())[0]['kernel'] == 'DRT'
    out[p + 'Rp'] = np.float64(inv.predict_Rp()) if single_drt else np.float64('nan')
    s_re, s_im = inv.predict_sigma(freq)
    out[p + 'sigma_train'] = np.concatenate((s_re, s_im))
    s_re, s_im = inv.predict_sigma(f_pred)
    out[p + 'sigma_pred'] = np.concatenate((s_re, s_im))
    out[p + 'score_chi_sq'] = np.float64(inv.score(freq, Z))
    out[p + 'score_r2_modulus'] = np.float64(inv.score(freq, Z, metric='r2', weights='modulus'))
    out[p + 'outlier_idx_z1'] = np.asarray(inv.check_outliers(freq, Z, threshold=1.0, use_existing_fit=True)).ravel()
    if fkw['mode'] == 'sample':
        out[p + 'Z_pred_p25'] = inv.predict_Z(f_pred, percentile=25)
    if fkw['mode'] == 'sample' and len(inv.distributions) == 1 and not single_drt:
        # single parallel distribution: percentile bands of the error structure off the training grid go through
        # predict_Z(percentile=...) with the parallel arithmetic Z = 1 / (A (x / scale)) (inversion.py:2712-2725)
        s_re, s_im = inv.predict_sigma(f_pred, percentile=60)
        out[p + 'sigma_pred_p60'] = np.concatenate((s_re, s_im))
    if fkw['mode'] == 'sample' and single_drt:
        out[p + 'Rp_p75'] = np.float64(inv.predict_Rp(percentile=75))
        s_re, s_im = inv.predict_sigma(f_pred, percentile=60)
        out[p + 'sigma_pred_p60'] = np.concatenate((s_re, s_im))
        s_re, s_im = inv.predict_sigma(freq, percentile=60)
        out[p + 'sigma_train_p60'] = np.concatenate((s_re, s_im))
    print(case, STATE['model'], 'R_inf %.5f Rp %.5f' % (inv.R_inf, out[p + 'Rp']), 'outliers(z>1):', len(out[p + 'outlier_idx_z1']))
dst = os.path.join(ROOT, 'tests', 'golden', 'postfit.npz')
np.savez_compressed(dst, **out)
print('wrote', dst, os.path.getsize(dst), 'bytes')
