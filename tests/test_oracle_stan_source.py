"""CPU: the oracle's hand-derived log-posterior and gradient (oracle/model.py, oracle/model_sp.py) against values
evaluated mechanically from the reference's own Stan source text.

tests/golden/stan_logdensity.npz was written by scripts/make_golden_stan_logdensity.py: a small interpreter of the Stan
subset those files use (scripts/stan_subset_interpreter.py) reads bayes_drt/stan_model_files/*_modelcode.txt in place
and evaluates log p + autograd gradient at three unconstrained points per program, for all nine programs the CUDA
engine implements, both constant sets ('optimize' / 'sample'), with and without the Jacobian.  The interpreter keeps
every normalising constant, Stan's `~` (and the oracle) drop the parameter-independent ones: log-densities are compared
as differences between points.  Points where Stan would reject (x_sum_raw < 0) must be -inf on both sides."""
import os

import numpy as np
import pytest

from oracle import model as omod, model_sp as osp

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'stan_logdensity.npz'))
FREQ, Z, BF = G['freq'], G['Z'], G['basis_freq']
TP = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'transmissive', 'dist_type': 'parallel', 'basis_freq': BF}
BP = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'blocking', 'dist_type': 'parallel', 'basis_freq': BF}
DRT = {'kernel': 'DRT', 'basis_freq': BF}
PROGRAMS = ['Series', 'Series_pos', 'Series_outliers', 'Series_pos_outliers', 'Parallel', 'Series-Parallel',
            'Series-Parallel_pos', 'Series-2Parallel', 'Series-2Parallel_pos']


def _oracle(program, mode):
    nonneg, outl = '_pos' in program, '_outliers' in program
    if program.startswith('Series-2Parallel'):
        return osp, osp.prep_series_2parallel(FREQ, Z, DRT, BP, dict(TP, x_scale=0.8), mode=mode, nonneg=nonneg)
    if program.startswith('Series-Parallel'):
        return osp, osp.prep_series_parallel(FREQ, Z, DRT, dict(TP, x_scale=0.8), mode=mode, nonneg=nonneg)
    if program == 'Parallel':
        return omod, omod.prep_parallel(FREQ, Z, TP, mode=mode)
    return omod, omod.prep_series(FREQ, Z, basis_freq=BF, mode=mode, nonneg=nonneg, outliers=outl)


@pytest.mark.parametrize('mode', ['optimize', 'sample'])
@pytest.mark.parametrize('program', PROGRAMS)
def test_logpost_matches_the_stan_source(program, mode):
    mod, d = _oracle(program, mode)
    key = f'{program}/{mode}'
    U = G[key + '/U']
    assert U.shape[1] == mod.n_params(d)
    for jac in (False, True):
        ref_lp, ref_g = G[key + f'/lp_jac{int(jac)}'], G[key + f'/grad_jac{int(jac)}']
        with np.errstate(all='ignore'):
            res = [mod.logpost(U[i], d, jacobian=jac) for i in range(3)]
        lp = np.array([r[0] for r in res])
        assert np.array_equal(np.isfinite(lp), np.isfinite(ref_lp)), (key, jac)  # the same points are rejected
        ok = np.nonzero(np.isfinite(ref_lp))[0]
        for i in ok[1:]:  # constants cancel in differences
            assert abs((lp[i] - lp[ok[0]]) - (ref_lp[i] - ref_lp[ok[0]])) <= 1e-11 * abs(ref_lp[i] - ref_lp[ok[0]]), (key, jac, i)
        for i in ok:
            g = res[i][1]
            assert np.max(np.abs(g - ref_g[i])) <= 1e-10 * np.max(np.abs(ref_g[i])), (key, jac, i)


@pytest.mark.parametrize('mode', ['optimize', 'sample'])
@pytest.mark.parametrize('program', PROGRAMS)
def test_constrain_matches_the_stan_source(program, mode):
    """The parameters and transformed parameters Stan reports (what the reference reads back, inversion.py:1229-1276)."""
    mod, d = _oracle(program, mode)
    key = f'{program}/{mode}'
    out = mod.constrain(G[key + '/U'][0], d)
    names = [k[len(key + '/tp/'):] for k in G.files if k.startswith(key + '/tp/')]
    assert {'Rinf', 'induc', 'sigma_tot', 'sigma_res', 'alpha_prop', 'alpha_re', 'alpha_im'} <= set(names)
    for nm in names:
        ref = G[key + '/tp/' + nm]
        assert np.max(np.abs(np.asarray(out[nm]) - ref)) <= 1e-12 * np.max(np.abs(ref)), (key, nm)
