"""GPU: batched bound-constrained QP and the hyper-lambda ridge loop against the oracle (exact active-set restatement
of the reference's cvxopt path).  North-star tolerance: ridge solutions within 1e-6 (relative to the largest
coefficient); QP KKT residual <= 1e-9."""
import numpy as np
import pytest
import torch

from helpers import load_spectrum
from oracle import ridge as oridge

pytestmark = pytest.mark.gpu
NAMES = ['ZARC_uniform_0.25', 'ZARC-RL_uniform_0.25', '2ZARC_uniform_0.25', 'Gerischer_uniform_0.25']


def test_qp_bound_random_problems():
    from bayes_drt_b200 import capi
    rng = np.random.RandomState(0)
    for n, B in ((7, 5), (40, 9), (103, 300)):
        A = rng.standard_normal((B, 2 * n, n))
        P = np.einsum('bij,bik->bjk', A, A) + 1e-3 * np.eye(n)
        q = rng.standard_normal((B, n)) * 3
        lb = np.where(rng.rand(n) < 0.5, 0.0, -0.7)
        x, kkt, iters = capi.qp_bound(torch.tensor(P), torch.tensor(q), torch.tensor(lb))
        x, kkt = x.cpu().numpy(), kkt.cpu().numpy()
        for b in range(0, B, max(1, B // 7)):
            xo, yo, F, it = oridge.qp_bound(P[b], q[b], lb)
            assert np.max(np.abs(x[b] - xo)) <= 1e-9 * max(1.0, np.max(np.abs(xo)))
        assert np.all(kkt <= 1e-9 * (1 + np.abs(q).max()))
        assert np.all(x >= lb - 1e-12)


@pytest.mark.parametrize('kw', [dict(), dict(preset='Huang'),
                                dict(penalty='integral', lambda_0=1, hl_beta=5, weights='modulus'),
                                dict(nonneg=False), dict(reg_ord=[0.2, 0.3, 0.5], L1_penalty=0.01),
                                dict(part='real'), dict(part='imag', weights='modulus'), dict(hl_fbeta=0.1),
                                dict(part='real', hl_fbeta=0.3, lambda_0=1e-3), dict(penalty='cholesky')])
@pytest.mark.parametrize('stop_rule', ['unchanged', 'nan'])
def test_ridge_fit_matches_oracle(kw, stop_rule):
    """Both stop rules: 'nan' is the reference's own code with an exact QP solver (all max_iter iterations), 'unchanged'
    (the default) lets the loop stop once the free coefficients have converged -- same iteration counts as the oracle."""
    from bayes_drt_b200 import Inverter
    freq = load_spectrum(NAMES[0])[0]
    Z = np.stack([load_spectrum(n)[1] for n in NAMES])
    inv = Inverter()
    inv.ridge_fit(freq, Z, stop_rule=stop_rule, **kw)
    coef = inv.distribution_fits['DRT']['coef'].cpu().numpy()
    for b in range(len(NAMES)):
        o = oridge.ridge_fit(freq, Z[b], stop_rule=stop_rule, **kw)
        scale = np.max(np.abs(o['coef']))
        assert inv._ridge_iters[b].item() == o['iters']
        assert np.max(np.abs(coef[b] - o['coef'])) <= 1e-6 * scale, (b, np.max(np.abs(coef[b] - o['coef'])) / scale)
        assert abs(inv.R_inf[b].item() - o['R_inf']) <= 1e-6 * max(abs(o['R_inf']), scale)
        assert abs(inv.inductance[b].item() - o['inductance']) <= 1e-6 * max(abs(o['inductance']), 1e-6)
        lam = inv.distribution_fits['DRT']['lambda_vectors'][b].cpu().numpy()
        assert np.allclose(lam, o['lam'], rtol=1e-5, atol=1e-12)


def test_ridge_single_spectrum_and_plain_ridge():
    """single-spectrum call returns reference shapes (numpy); hyper_lambda=False is one QP at lambda_0."""
    from bayes_drt_b200 import Inverter
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    inv = Inverter()
    with pytest.warns(UserWarning):  # 'did not converge within 20 iterations' (exact zeros -> NaN stop test)
        inv.ridge_fit(freq, Z, stop_rule='nan')
    assert isinstance(inv.distribution_fits['DRT']['coef'], np.ndarray) and inv.distribution_fits['DRT']['coef'].shape == (101,)
    assert inv.fit_type == 'ridge' and abs(inv.R_inf - 0.9914) < 2e-4  # SURVEY section 7 1b sanity value
    rp = inv.predict_Rp()
    assert abs(rp - 1.0204) < 2e-4
    inv.ridge_fit(freq, Z, hyper_lambda=False, lambda_0=1e-2)
    p = oridge.prep(freq, Z)
    G0 = p['WA_re'].T @ p['WA_re'] + p['WA_im'].T @ p['WA_im']
    q = -p['WA_re'].T @ p['WZ_re'] - p['WA_im'].T @ p['WZ_im']
    xo, _, _, _ = oridge.qp_bound(G0 + 1e-2 * p['Pen'][2], q, np.zeros(len(q)))
    assert np.max(np.abs(inv.distribution_fits['DRT']['scaled_coef'] - xo)) <= 1e-8 * np.max(np.abs(xo))


def test_ridge_unsupported_options_are_loud():
    from bayes_drt_b200 import Inverter
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    inv = Inverter()
    for kw in (dict(hl_fbeta=0.1, penalty='integral', hl_beta=2.5), dict(hyper_weights=True, hyper_lambda=False),
               dict(hl_solution='lm'), dict(dZ=True)):
        with pytest.raises(NotImplementedError):
            inv.ridge_fit(freq, Z, **kw)
    with pytest.raises(ValueError):
        inv.ridge_fit(freq, Z, hl_beta=0.5)
    with pytest.raises(ValueError):
        inv.ridge_fit(freq, Z, preset='nope')
    with pytest.raises(ValueError):
        inv.ridge_fit(freq, Z, part='modulus')


def test_ridge_reim_cross_validation_and_ciucci_preset():
    """lambda_0='cv' (Inverter.ridge_ReImCV, inversion.py:902-944) and preset 'Ciucci' against the oracle: the
    cross-validation table, the selected lambda_0 of every spectrum of a batch, and the final fit."""
    import warnings
    from bayes_drt_b200 import Inverter
    freq = load_spectrum(NAMES[0])[0]
    Z = np.stack([load_spectrum(n)[1] for n in NAMES[:3]])
    grid = np.logspace(-6, 0, 7)
    inv = Inverter()
    inv.ridge_fit(freq, Z, lambda_0='cv', cv_lambdas=grid, stop_rule='nan')
    coef = inv.distribution_fits['DRT']['coef'].cpu().numpy()
    for b in range(3):
        o = oridge.ridge_fit(freq, Z[b], lambda_0='cv', cv_lambdas=grid)
        assert inv._cv_lambda_0[b] == o['lambda_0']
        tab = o['cv_result']
        zs2 = o['Z_scale'] ** 2
        for j, key in ((1, 'recv'), (2, 'imcv'), (3, 'totcv')):
            assert np.allclose(inv.cv_result[key][b].cpu().numpy(), tab[:, j] * zs2, rtol=1e-5), key
        assert np.max(np.abs(coef[b] - o['coef'])) <= 1e-6 * np.max(np.abs(o['coef']))
        assert abs(inv.R_inf[b].item() - o['R_inf']) <= 1e-6 * abs(o['R_inf'])
    # single spectrum, full default grid, preset 'Ciucci' (discrete penalty, hl_fbeta = 0.1)
    one = Inverter()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        one.ridge_fit(freq, Z[0], preset='Ciucci', stop_rule='nan')
    o = oridge.ridge_fit(freq, Z[0], preset='Ciucci')
    assert list(one.cv_result.columns) == ['lambda', 'recv', 'imcv', 'totcv'] and len(one.cv_result) == 31
    assert one._cv_lambda_0[0] == o['lambda_0']
    c1 = one.distribution_fits['DRT']['coef']
    assert np.max(np.abs(c1 - o['coef'])) <= 1e-6 * np.max(np.abs(o['coef']))
    assert abs(one.predict_Rp() - 1.0) < 0.05


def test_ridge_default_stop_rule_converges_and_large_nf():
    """Default stop rule: the default fit of the reference spectrum stops early with the converged flag set (no
    'did not converge' warning), close to the max_iter = 20 solution; and a spectrum with many more frequencies than
    basis functions (Nf = 161 > K + 2 = 103: 20 points per decade) is fitted like any other."""
    import warnings
    from bayes_drt_b200 import Inverter
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    inv = Inverter()
    with warnings.catch_warnings():
        warnings.simplefilter('error')
        inv.ridge_fit(freq, Z)
    assert bool(inv._ridge_converged[0]) and inv._ridge_iters[0].item() < 20
    o = oridge.ridge_fit(freq, Z, stop_rule='unchanged')
    assert o['converged'] and inv._ridge_iters[0].item() == o['iters']
    assert np.max(np.abs(inv.distribution_fits['DRT']['coef'] - o['coef'])) <= 1e-6 * np.max(np.abs(o['coef']))
    full = oridge.ridge_fit(freq, Z, stop_rule='nan')  # all 20 iterations
    assert np.max(np.abs(o['coef'] - full['coef'])) <= 0.05 * np.max(np.abs(full['coef']))
    # dense frequency grid, default basis (K = 101): Nf > K + 2
    f2 = np.logspace(6, -2, 161)
    w = 2 * np.pi * f2
    rng = np.random.RandomState(0)
    Z2 = 1.0 + 1.0 / (1 + (1j * w * 1e-2) ** 0.8) + 0.002 * (rng.standard_normal(161) + 1j * rng.standard_normal(161))
    inv2 = Inverter()
    inv2.ridge_fit(f2, Z2)
    o2 = oridge.ridge_fit(f2, Z2, stop_rule='unchanged')
    assert inv2._ridge_iters[0].item() == o2['iters']
    assert np.max(np.abs(inv2.distribution_fits['DRT']['coef'] - o2['coef'])) <= 1e-6 * np.max(np.abs(o2['coef']))
    assert abs(inv2.R_inf - o2['R_inf']) <= 1e-6 * abs(o2['R_inf'])


def test_ridge_per_spectrum_grids():
    """ridge_fit with one frequency grid per spectrum ([B, Nf]): every row equals the single-grid fit of that row."""
    from bayes_drt_b200 import Inverter
    rng = np.random.RandomState(1)
    B, Nf = 5, 81
    shift = rng.uniform(0, 1, B)
    f = 10.0 ** (6 - shift[:, None] - np.arange(Nf)[None, :] / 10.0)
    w = 2 * np.pi * f
    Z = 0.8 + 1.2 / (1 + (1j * w * 10.0 ** (-3 + shift[:, None])) ** 0.85) + \
        0.003 * (rng.standard_normal((B, Nf)) + 1j * rng.standard_normal((B, Nf)))
    for kw in (dict(), dict(preset='Huang')):
        inv = Inverter()
        inv.ridge_fit(f, Z, **kw)
        coef = inv.distribution_fits['DRT']['coef'].cpu().numpy()
        for b in range(B):
            one = Inverter()
            one.ridge_fit(f[b], Z[b], **kw)
            sc = np.max(np.abs(one.distribution_fits['DRT']['coef']))
            assert np.max(np.abs(coef[b] - one.distribution_fits['DRT']['coef'])) <= 1e-7 * sc
            assert abs(inv.R_inf[b].item() - one.R_inf) <= 1e-7 * abs(one.R_inf)
            assert inv._ridge_iters[b].item() == one._ridge_iters[0].item()
