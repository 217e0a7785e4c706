"""CPU: the oracle's hand-derived log-posterior / gradient (oracle/model.py) against an independent line-by-line
transcription of the Stan programs differentiated by torch autograd (oracle/stan_literal.py), for every model variant
the CUDA engine implements, with and without the Jacobian; plus the loose paper goldens:
the oracle's Stan-semantics L-BFGS on the reference's simulated spectra lands on the published MAP DRT."""
import numpy as np
import pytest
import torch

from helpers import load_spectrum
from oracle import lbfgs as olb, model as omod
from oracle.stan_literal import logpost_literal


@pytest.mark.parametrize('nonneg,outliers', [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize('mode', ['optimize', 'sample'])
def test_logpost_matches_literal_autograd(nonneg, outliers, mode):
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    d = omod.prep_series(freq, Z, mode=mode, nonneg=nonneg, outliers=outliers)
    D = omod.n_params(d)
    assert D == 2 * d['K'] + 9 + (2 * d['Nf'] if outliers else 0)
    rng = np.random.RandomState(5)
    for jac in (False, True):
        for scale in (0.3, 1.0):
            u = rng.uniform(-scale, scale, D)
            lp, g = omod.logpost(u, d, jacobian=jac)
            ut = torch.tensor(u, requires_grad=True)
            lt = logpost_literal(ut, d, jacobian=jac)
            lt.backward()
            assert abs(lp - lt.item()) <= 1e-12 * abs(lt.item())
            gt = ut.grad.numpy()
            assert np.max(np.abs(g - gt)) <= 1e-10 * np.max(np.abs(gt))


def test_constrain_matches_model_definitions():
    freq, Z = load_spectrum('ZARC-RL_uniform_0.25')
    d = omod.prep_series(freq, Z, mode='optimize', outliers=True)
    u = np.random.RandomState(0).uniform(-1, 1, omod.n_params(d))
    c = omod.constrain(u, d)
    Nf = d['Nf']
    assert c['Rinf'] == 100 * np.exp(u[0]) and c['induc'] == np.exp(u[1]) * d['induc_scale']
    zh = d['A'] @ c['x']
    zh[:Nf] += c['Rinf']
    zh[Nf:] += c['induc'] * 2 * np.pi * d['freq']
    assert np.allclose(zh, c['Z_hat'], rtol=1e-14)
    # Series_outliers_modelcode.txt:45-51
    var = d['sigma_min'] ** 2 + c['sigma_res'] ** 2 + (c['alpha_prop'] * zh) ** 2 \
        + np.tile((c['alpha_re'] * zh[:Nf]) ** 2 + (c['alpha_im'] * zh[Nf:]) ** 2 + c['sigma_out'] ** 2, 2)
    assert np.allclose(np.sqrt(var), c['sigma_tot'], rtol=1e-14)


def test_preprocessing_follows_reference_defaults():
    """inversion.py:2191-2209 (default basis grid and epsilon), :2437-2441 (Z scale), :1725-1737 (mode constants)."""
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    d = omod.prep_series(freq, Z, mode='optimize')
    assert d['Nf'] == 81 and d['K'] == 101
    assert abs(d['epsilon'] - 4.342944819) < 1e-8
    assert np.isclose(d['tau'][0], 1 / (2 * np.pi * 1e6) / 10) and np.isclose(d['tau'][-1], 1 / (2 * np.pi * 1e-2) * 10)
    assert np.isclose(d['Z_scale'], np.std(np.abs(Z)))
    assert (d['ups_alpha'], d['ups_beta']) == (0.05, 0.1)
    # L0[n, n] = l0 * exp(0) ; banded Toeplitz with closed-form taps at the default epsilon
    assert np.isclose(d['L0'][50, 50], 0.36) and np.isclose(d['L0'][50, 51], 0.36 * np.exp(-1), rtol=1e-9)
    assert np.isclose(d['L2'][50, 50], 0.12 * -2 * d['epsilon'] ** 2)
    s = omod.prep_series(freq, Z, mode='sample')
    assert (s['ups_alpha'], s['ups_beta']) == (1.0, 0.1) and np.isclose(s['L2'][50, 50], 0.75 * -2 * s['epsilon'] ** 2)


@pytest.mark.parametrize('name,tol', [('ZARC-RL_uniform_0.25', 0.02), ('ZARC_uniform_0.25', 0.03)])
def test_oracle_map_lands_on_paper_drt(name, tol):
    """Loose golden: code_EchemActa/map_results/Gout_<name>.csv (legacy code, same Stan maths; SURVEY section 4 measured
    0.5-1.2 % of peak between a converged restatement and the published curve)."""
    g = np.load(__import__('os').path.join(__import__('helpers').GOLD, 'spectra.npz'))
    freq, Z = g[name + '/freq'], g[name + '/Z']
    d = omod.prep_series(freq, Z, mode='optimize')

    def f(u):
        with np.errstate(all='ignore'):
            lp, gr = omod.logpost(u, d)
        if not np.isfinite(lp) or not np.all(np.isfinite(gr)):
            return None
        return -lp, -gr
    # a sane deterministic start (x = 0.01, raw scales = exp(-1)): the default model has a unique optimum
    u0 = np.full(omod.n_params(d), -1.0)
    u0[2:2 + d['K']] = 0.01
    r = olb.minimize(f, u0, max_iter=6000)
    c = omod.constrain(r['x'], d)
    tau_e = g[name + '/map_tau']
    phi = np.exp(-(d['epsilon'] * np.log(tau_e[:, None] / d['tau'][None, :])) ** 2)
    gamma = phi @ (c['x'] * d['Z_scale'])
    gold = g[name + '/map_gamma']
    assert np.max(np.abs(gamma - gold)) <= tol * np.max(gold), np.max(np.abs(gamma - gold)) / np.max(gold)
    assert abs(c['Rinf'] * d['Z_scale'] - 1.0) < 0.02


# ---------------------------------------------------------------------------------------------------------------------
# Series-Parallel (DRT + transmissive planar DDT), the paper's two-distribution shape (Run fits.ipynb cell 20)
# ---------------------------------------------------------------------------------------------------------------------
def _sp_data(mode, nonneg, Nf=41, K=33):
    from oracle import model_sp as osp
    rng = np.random.RandomState(3)
    freq = np.logspace(5, -1, Nf)
    bf = np.logspace(5.5, -1.5, K)
    # a DRT arc in series with a transmissive diffusion element
    tau0, R1, n, Rd, td = 1e-3, 1.0, 0.8, 0.7, 0.3
    w = 2 * np.pi * freq
    Zd = Rd * np.tanh(np.sqrt(1j * w * td)) / np.sqrt(1j * w * td)
    Z = 0.5 + R1 / (1 + (1j * w * tau0) ** n) + Zd
    Z = Z + 0.002 * (rng.standard_normal(Nf) + 1j * rng.standard_normal(Nf))
    ser = {'kernel': 'DRT', 'basis_freq': bf}
    par = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'transmissive', 'dist_type': 'parallel', 'basis_freq': bf,
           'x_scale': 0.8}
    return osp.prep_series_parallel(freq, Z, ser, par, mode=mode, nonneg=nonneg)


@pytest.mark.parametrize('nonneg', [True, False])
@pytest.mark.parametrize('mode', ['optimize', 'sample'])
def test_series_parallel_logpost_matches_literal_autograd(nonneg, mode):
    from oracle import model_sp as osp
    from oracle.stan_literal import logpost_literal_sp
    d = _sp_data(mode, nonneg)
    D = osp.n_params(d)
    assert D == 2 * (d['Ks'] + d['Kp']) + 12
    rng = np.random.RandomState(9)
    for jac in (False, True):
        for scale in (0.3, 1.0):
            u = rng.uniform(-scale, scale, D)
            if not nonneg:
                u[2:2 + d['Ks']] = np.abs(u[2:2 + d['Ks']])  # keep x_sum_raw >= 0 (Stan rejects otherwise)
            lp, g = osp.logpost(u, d, jacobian=jac)
            ut = torch.tensor(u, requires_grad=True)
            lt = logpost_literal_sp(ut, d, jacobian=jac)
            lt.backward()
            assert abs(lp - lt.item()) <= 1e-12 * abs(lt.item())
            gt = ut.grad.numpy()
            assert np.max(np.abs(g - gt)) <= 1e-10 * np.max(np.abs(gt))
    if not nonneg:  # validity check of the transformed parameter x_sum_raw
        u = rng.uniform(-0.3, 0.3, D)
        u[2:2 + d['Ks']] = -5.0
        assert osp.logpost(u, d)[0] == -np.inf
        assert logpost_literal_sp(torch.tensor(u), d).item() == -np.inf


def test_series_parallel_constants_follow_reference():
    d = _sp_data('optimize', True)
    assert d['x_sum_invscale'] == 0.0 and d['xp_scale'] == 0.8 and (d['ups_alpha'], d['ups_beta']) == (0.05, 0.1)
    k = d['Ks'] // 2
    # inversion.py:1921-1926: L0s x 0.36, L0p x 0.54
    assert np.isclose(d['Ls'][0][k, k], 0.36) and np.isclose(d['Lp'][0][k, k], 0.54)
    s = _sp_data('sample', True)
    assert s['x_sum_invscale'] == 1.0 and np.isclose(s['Lp'][0][k, k], 1.0)


# ---------------------------------------------------------------------------------------------------------------------
# Parallel (single DDT): Z_hat = 1 / (A x) + offsets
# ---------------------------------------------------------------------------------------------------------------------
def _par_data(mode, bc='transmissive', Nf=41, K=37):
    rng = np.random.RandomState(6)
    freq = np.logspace(4, -2, Nf)
    w = 2 * np.pi * freq
    x = np.sqrt(1j * w * 0.5)
    Zd = 0.8 * (np.tanh(x) / x if bc == 'transmissive' else 1 / (x * np.tanh(x)))
    Z = 0.3 + Zd + 0.002 * (rng.standard_normal(Nf) + 1j * rng.standard_normal(Nf))
    info = {'kernel': 'DDT', 'dist_type': 'parallel', 'symmetry': 'planar', 'bc': bc, 'basis_freq': np.logspace(4.5, -2.5, K)}
    return omod.prep_parallel(freq, Z, info, mode=mode), Z


@pytest.mark.parametrize('mode', ['optimize', 'sample'])
@pytest.mark.parametrize('bc', ['transmissive', 'blocking'])
def test_parallel_logpost_matches_literal_autograd(mode, bc):
    d, Z = _par_data(mode, bc)
    assert d['parallel'] and d['pos']
    # inversion.py:2417-2434: scaled admittance std 14 (transmissive) / 2.4 (blocking) x sqrt(Nf / 81)
    assert np.isclose(np.std(np.abs(d['Z_scale'] / Z)), (14.0 if bc == 'transmissive' else 2.4) * np.sqrt(41 / 81))
    D = omod.n_params(d)
    rng = np.random.RandomState(10)
    for jac in (False, True):
        u = rng.uniform(-1, 1, D)
        lp, g = omod.logpost(u, d, jacobian=jac)
        ut = torch.tensor(u, requires_grad=True)
        lt = logpost_literal(ut, d, jacobian=jac)
        lt.backward()
        assert abs(lp - lt.item()) <= 1e-12 * abs(lt.item())
        assert np.max(np.abs(g - ut.grad.numpy())) <= 1e-10 * np.max(np.abs(ut.grad.numpy()))
    c = omod.constrain(u, d)
    Y = d['A'] @ c['x']
    zc = 1 / (Y[:41] + 1j * Y[41:]) + c['Rinf'] + 1j * c['induc'] * 2 * np.pi * d['freq']
    assert np.allclose(c['Z_hat'], np.r_[zc.real, zc.imag], rtol=1e-13)


@pytest.mark.parametrize('mode', ['optimize', 'sample'])
def test_series_2parallel_logpost_matches_literal_autograd(mode):
    """Series-2Parallel_pos: DRT + transmissive DDT + blocking DDT (the parallel distributions in sorted-name order)."""
    from oracle import model_sp as osp
    from oracle.stan_literal import logpost_literal_s2p
    rng = np.random.RandomState(12)
    Nf, K = 41, 29
    freq = np.logspace(5, -1, Nf)
    bf = np.logspace(5.5, -1.5, K)
    w = 2 * np.pi * freq
    Z = 0.5 + 1.0 / (1 + (1j * w * 1e-3) ** 0.8) + 0.7 * np.tanh(np.sqrt(1j * w * 0.3)) / np.sqrt(1j * w * 0.3)
    Z = Z + 0.002 * (rng.standard_normal(Nf) + 1j * rng.standard_normal(Nf))
    ser = {'kernel': 'DRT', 'basis_freq': bf}
    p1 = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'transmissive', 'basis_freq': bf, 'x_scale': 0.8}
    p2 = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'blocking', 'basis_freq': bf[:-4]}
    d = osp.prep_series_2parallel(freq, Z, ser, p1, p2, mode=mode)
    D = osp.n_params(d)
    assert D == 2 * (K + K + K - 4) + 15 and d['x_sum_invscale'] == (0.1 if mode == 'sample' else 0.0)
    for jac in (False, True):
        u = rng.uniform(-1, 1, D)
        lp, g = osp.logpost(u, d, jacobian=jac)
        ut = torch.tensor(u, requires_grad=True)
        lt = logpost_literal_s2p(ut, d, jacobian=jac)
        lt.backward()
        assert abs(lp - lt.item()) <= 1e-12 * abs(lt.item())
        assert np.max(np.abs(g - ut.grad.numpy())) <= 1e-10 * np.max(np.abs(ut.grad.numpy()))
    c = osp.constrain(u, d)
    assert c['xp1'].shape == (K,) and c['xp2'].shape == (K - 4,) and np.isclose(c['xp1'][0], 0.8 * np.exp(u[2 + K]))
