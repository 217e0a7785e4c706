"""CPU: the oracle's data preparation against the Stan input of the reference's own Inverter.fit.

tests/golden/stan_data.npz was produced by scripts/make_golden_stan_data.py, which imports bayes_drt/inversion.py
unmodified and records what it hands to StanModel.optimizing / .sampling (model file name, data dict, init) for fifteen
configurations.  Matrices are compared through random projections (they are pinned entry by entry in matrices.npz)."""
import os

import numpy as np
import pytest

from oracle import model as omod, model_sp as osp, ridge as oridge

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'stan_data.npz'))
FREQ, Z = G['freq'], G['Z']
BF = np.logspace(6, -2, 81)
TP = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'transmissive', 'dist_type': 'parallel', 'basis_freq': BF}
BP = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'blocking', 'dist_type': 'parallel', 'basis_freq': BF}
DRT = {'kernel': 'DRT', 'basis_freq': BF}


def _proj(case, key, M, rtol=1e-11):
    M = np.asarray(M, dtype=np.float64)
    r = np.random.RandomState(M.shape[0] * 1000 + M.shape[1])
    v, u = r.standard_normal(M.shape[1]), r.standard_normal(M.shape[0])
    a, b = G[f'{case}/{key}@v'], G[f'{case}/u@{key}']
    assert np.max(np.abs(M @ v - a)) <= rtol * np.max(np.abs(a)), (case, key)
    assert np.max(np.abs(u @ M - b)) <= rtol * np.max(np.abs(b)), (case, key)


def _scalars(case, d, keys):
    for k_ref, k in keys.items():
        assert float(G[f'{case}/{k_ref}']) == pytest.approx(float(d[k]), rel=1e-14), (case, k_ref)


SERIES = {
    'series_opt': dict(mode='optimize'),
    'series_sample': dict(mode='sample'),
    'series_pos_opt': dict(mode='optimize', nonneg=True),
    'series_out_opt': dict(mode='optimize', outliers=True),
    'series_pos_out_sample': dict(mode='sample', nonneg=True, outliers=True, outlier_lambda=5, sigma_min=0.001,
                                  inductance_scale=2),
    'series_basis_eq_freq': dict(mode='optimize', basis_freq=FREQ),
    'series_noscale': dict(mode='optimize', scale_Z=False),
}


@pytest.mark.parametrize('case', sorted(SERIES))
def test_series_stan_data(case):
    kw = SERIES[case]
    d = omod.prep_series(FREQ, Z, **kw)
    name = 'Series' + ('_pos' if kw.get('nonneg') else '') + ('_outliers' if kw.get('outliers') else '') + '_StanModel.pkl'
    assert str(G[f'{case}/model']) == name  # Inverter._get_stan_model, inversion.py:1576-1610
    assert int(G[f'{case}/K']) == d['K'] and np.array_equal(G[f'{case}/freq'], d['freq'])
    assert int(G[f'{case}/N']) == (d['Nf'] if kw.get('outliers') else 2 * d['Nf'])  # outlier programs: N = Nf (:1208-1211)
    assert float(G[f'{case}/Z_scale']) == pytest.approx(d['Z_scale'], rel=1e-14)
    assert np.max(np.abs(G[f'{case}/Z'] - d['Z'])) <= 1e-14 * np.max(np.abs(d['Z']))
    for k in ('A', 'L0', 'L1', 'L2'):
        _proj(case, k, d[k])
    _scalars(case, d, dict(sigma_min='sigma_min', ups_alpha='ups_alpha', ups_beta='ups_beta', induc_scale='induc_scale'))
    if kw.get('outliers'):
        _scalars(case, d, dict(sigma_out_lambda='sigma_out_lambda', sigma_out_alpha='sigma_out_alpha',
                               sigma_out_beta='sigma_out_beta'))
    if kw['mode'] == 'optimize':
        assert int(G[f'{case}/iter']) == 50000 and int(G[f'{case}/seed']) == 1234 and str(G[f'{case}/init']) == 'random'
    else:  # inversion.py:1218-1221
        assert int(G[f'{case}/warmup']) == 200 and int(G[f'{case}/iter']) == 400 and int(G[f'{case}/chains']) == 2
        assert list(G[f'{case}/control']) == [0.9, 10]


@pytest.mark.parametrize('case,mode', [('sp_opt', 'optimize'), ('sp_sample', 'sample')])
def test_series_parallel_stan_data(case, mode):
    d = osp.prep_series_parallel(FREQ, Z, DRT, dict(TP, x_scale=0.8), mode=mode, nonneg=True)
    assert str(G[f'{case}/model']) == 'Series-Parallel_pos_StanModel.pkl'
    assert int(G[f'{case}/Ks']) == d['Ks'] and int(G[f'{case}/Kp']) == d['Kp'] and int(G[f'{case}/N']) == 2 * d['Nf']
    assert np.max(np.abs(G[f'{case}/Z'] - d['Z'])) <= 1e-14 * np.max(np.abs(d['Z']))
    _proj(case, 'As', d['As'])
    _proj(case, 'Ap', d['Ap'])
    for o in range(3):
        _proj(case, f'L{o}s', d['Ls'][o])
        _proj(case, f'L{o}p', d['Lp'][o])
    _scalars(case, d, dict(sigma_min='sigma_min', ups_alpha='ups_alpha', ups_beta='ups_beta', induc_scale='induc_scale',
                           x_sum_invscale='x_sum_invscale', xp_scale='xp_scale'))


@pytest.mark.parametrize('case,info,mode', [('parallel_tp_opt', TP, 'optimize'), ('parallel_bp_sample', BP, 'sample')])
def test_parallel_stan_data(case, info, mode):
    d = omod.prep_parallel(FREQ, Z, info, mode=mode)
    assert str(G[f'{case}/model']) == 'Parallel_StanModel.pkl'
    assert float(G[f'{case}/Z_scale']) == pytest.approx(d['Z_scale'], rel=1e-13)  # admittance scaling, :2417-2434
    assert np.max(np.abs(G[f'{case}/Z'] - d['Z'])) <= 1e-13 * np.max(np.abs(d['Z']))
    for k in ('A', 'L0', 'L1', 'L2'):
        _proj(case, k, d[k])
    _scalars(case, d, dict(sigma_min='sigma_min', ups_alpha='ups_alpha', ups_beta='ups_beta', induc_scale='induc_scale'))


@pytest.mark.parametrize('case,mode', [('s2p_opt', 'optimize'), ('s2p_sample', 'sample')])
def test_series_2parallel_stan_data(case, mode):
    # the reference orders the parallel distributions by name: 'BP-DDT' < 'TP-DDT'
    d = osp.prep_series_2parallel(FREQ, Z, DRT, BP, dict(TP, x_scale=0.8), mode=mode, nonneg=True)
    assert str(G[f'{case}/model']) == 'Series-2Parallel_pos_StanModel.pkl'
    _proj(case, 'As', d['As'])
    _proj(case, 'Ap1', d['Ap'])
    _proj(case, 'Ap2', d['Ap2'])
    for o in range(3):
        _proj(case, f'L{o}s', d['Ls'][o])
        _proj(case, f'L{o}p1', d['Lp'][o])
        _proj(case, f'L{o}p2', d['Lp2'][o])
    _scalars(case, d, dict(sigma_min='sigma_min', ups_alpha='ups_alpha', ups_beta='ups_beta', induc_scale='induc_scale',
                           x_sum_invscale='x_sum_invscale', xp1_scale='xp_scale', xp2_scale='xp2_scale'))


@pytest.mark.parametrize('case,nonneg,outliers', [('series_ridge_init', False, False),
                                                  ('series_pos_out_ridge_init', True, True)])
def test_init_from_ridge(case, nonneg, outliers):
    """Inverter._get_init_from_ridge (inversion.py:1616-1682): an under-fitted ridge solution in scaled units.  The
    ridge fit keeps its own default nonneg=True whatever the Stan model's constraint is (:1643 does not forward it)."""
    r = oridge.ridge_fit(FREQ, Z, penalty='integral', lambda_0=1, hl_beta=5, weights='modulus')
    zs = r['Z_scale']
    x = G[f'{case}/init/x']
    assert np.max(np.abs(x - r['coef'] / zs)) <= 1e-9 * np.max(np.abs(x))
    assert float(G[f'{case}/init/Rinf']) == pytest.approx(r['R_inf'] / zs, rel=1e-9)
    assert float(G[f'{case}/init/Rinf_raw']) == pytest.approx(r['R_inf'] / zs / 100, rel=1e-9)
    induc = r['inductance'] / zs
    assert float(G[f'{case}/init/induc']) == pytest.approx(induc if induc > 0 else 1e-10, rel=1e-9)
    if outliers:
        assert np.array_equal(G[f'{case}/init/sigma_out_raw'], np.full(len(FREQ), 0.1))
