"""GPU: the Inverter facade end to end (the call a user of the reference makes) -- single spectrum in reference shapes,
batches, both modes, the outlier model, the two-distribution model, post-fit queries, and loud errors for what the CUDA
path does not implement.  Loose goldens: the paper's published MAP / HMC DRT curves."""
import numpy as np
import pytest
import torch

from helpers import GOLD, load_spectrum, sp_dists, sp_spectrum

pytestmark = pytest.mark.gpu


def _gold():
    import os
    return np.load(os.path.join(GOLD, 'spectra.npz'))


def test_fit_map_single_spectrum_reference_shapes():
    from bayes_drt_b200 import Inverter
    name = 'ZARC-RL_uniform_0.25'
    freq, Z = load_spectrum(name)
    inv = Inverter()
    inv.fit(freq, Z, mode='optimize')
    assert inv.fit_type == 'map' and inv.stan_model_name == 'Series_StanModel.pkl'
    coef = inv.distribution_fits['DRT']['coef']
    assert isinstance(coef, np.ndarray) and coef.shape == (101,)
    assert isinstance(inv.R_inf, float) and isinstance(inv.inductance, float)
    assert inv.error_fit['sigma_tot'].shape == (162,) and inv.error_fit['sigma_res'] > 0
    assert abs(inv.R_inf - 1.0) < 0.02
    g = _gold()
    gamma = inv.predict_distribution('DRT', eval_tau=g[name + '/map_tau'])
    gold = g[name + '/map_gamma']
    assert np.max(np.abs(gamma - gold)) <= 0.03 * gold.max()  # paper curve (legacy code, Stan's own termination)
    Zp = inv.predict_Z(freq)
    assert Zp.shape == (81,) and np.max(np.abs(Zp - Z)) < 0.02
    assert abs(inv.predict_Rp() - 0.8) < 0.03  # truth R_p = 0.8 for ZARC-RL
    s_re, s_im = inv.predict_sigma()
    assert s_re.shape == (81,) and (s_re > 0).all() and (s_im > 0).all()
    # frequencies in ascending order give the same fit (the reference sorts, inversion.py:2138-2141)
    inv2 = Inverter()
    inv2.fit(freq[::-1].copy(), Z[::-1].copy(), mode='optimize')
    assert np.allclose(inv2.distribution_fits['DRT']['coef'], coef, rtol=0, atol=1e-12)


def test_fit_map_batch_matches_single_and_polish():
    from bayes_drt_b200 import Inverter
    names = ['ZARC_uniform_0.25', 'ZARC-RL_uniform_0.25', '2ZARC_uniform_0.25']
    freq = load_spectrum(names[0])[0]
    Zb = np.stack([load_spectrum(n)[1] for n in names])
    inv = Inverter()
    inv.fit(freq, torch.tensor(Zb), mode='optimize', polish=True)
    coef = inv.distribution_fits['DRT']['coef']
    assert torch.is_tensor(coef) and coef.is_cuda and tuple(coef.shape) == (3, 101)
    assert tuple(inv.R_inf.shape) == (3,) and tuple(inv.error_fit['sigma_tot'].shape) == (3, 162)
    assert (inv._opt_result['gnorm'] < 1e-6).all()
    # spectrum 1 fitted alone with the same global index -> identical init -> identical result
    one = Inverter()
    one.fit(freq, Zb[1], mode='optimize', polish=True, spectrum_offset=1)
    assert np.allclose(one.distribution_fits['DRT']['coef'], coef[1].cpu().numpy(), rtol=0, atol=1e-9)
    rp = inv.predict_Rp()
    assert tuple(rp.shape) == (3,) and abs(rp[1].item() - 0.8) < 0.03


def test_fit_sample_percentiles_and_outlier_model():
    from bayes_drt_b200 import Inverter
    name = 'ZARC-RL_uniform_0.25'
    freq, Z = load_spectrum(name)
    inv = Inverter()
    inv.fit(freq, Z, mode='sample', chains=4, warmup=200, samples=100, nonneg=True, outliers=True)
    assert inv.fit_type == 'bayes' and inv.stan_model_name == 'Series_pos_outliers_StanModel.pkl'
    assert inv.error_fit['sigma_out'].shape == (81,)
    lo, mid, hi = (inv.coef_percentile('DRT', p) for p in (2.5, 50, 97.5))
    assert lo.shape == (101,) and (lo <= mid).all() and (mid <= hi).all() and (lo >= 0).all()
    g = _gold()
    tau = g[name + '/bayes_tau']
    glo = inv.predict_distribution('DRT', eval_tau=tau, percentile=2.5)
    ghi = inv.predict_distribution('DRT', eval_tau=tau, percentile=97.5)
    assert (glo <= ghi + 1e-12).all()
    # the paper's HMC curve was made with the default model (Series: nonneg=False, no outlier terms)
    ref = Inverter()
    ref.fit(freq, Z, mode='sample', chains=4, warmup=200, samples=200)
    gm = ref.predict_distribution('DRT', eval_tau=tau)
    gold = g[name + '/bayes_gamma']
    assert np.max(np.abs(gm - gold)) <= 0.06 * gold.max()  # SURVEY appendix A measured 2.2 % with 400 draws
    blo, bhi = (ref.predict_distribution('DRT', eval_tau=tau, percentile=p) for p in (2.5, 97.5))
    assert np.max(np.abs(blo - g[name + '/bayes_gamma_lo'])) <= 0.12 * gold.max()
    assert np.max(np.abs(bhi - g[name + '/bayes_gamma_hi'])) <= 0.12 * gold.max()
    # R_p = 0.8 for ZARC-RL (its inductive loop needs negative DRT values, so the default model is the one to check)
    rp_lo, rp, rp_hi = ref.predict_Rp(percentile=2.5), ref.predict_Rp(), ref.predict_Rp(percentile=97.5)
    assert rp_lo < rp < rp_hi and abs(rp - 0.8) < 0.03 and rp_hi - rp_lo < 0.1
    assert inv.predict_Rp(percentile=2.5) < inv.predict_Rp() < inv.predict_Rp(percentile=97.5)
    Zlo = inv.predict_Z(freq, percentile=2.5)
    Zhi = inv.predict_Z(freq, percentile=97.5)
    assert (Zlo.real <= Zhi.real).all()
    assert inv._sample_stats['n_leapfrog'].sum().item() > 0


def test_fit_series_parallel_distributions():
    from bayes_drt_b200 import Inverter
    freq = np.logspace(6, -2, 81)
    Z = np.stack([sp_spectrum(freq, seed=s, td=0.2 * (s + 1)) for s in range(2)])
    ser, par = sp_dists(np.logspace(6, -2, 81), np.logspace(6, -2, 81))
    inv = Inverter(distributions={'DRT': ser, 'TP-DDT': par})
    inv.fit(freq, torch.tensor(Z), mode='optimize', nonneg=True, max_iter=5000)
    assert inv.stan_model_name == 'Series-Parallel_pos_StanModel.pkl'
    cs, cp = inv.distribution_fits['DRT']['coef'], inv.distribution_fits['TP-DDT']['coef']
    assert tuple(cs.shape) == (2, 81) and tuple(cp.shape) == (2, 81) and (cs >= 0).all() and (cp >= 0).all()
    Zp = inv.predict_Z(freq)
    assert tuple(Zp.shape) == (2, 81)
    assert (Zp.cpu() - torch.tensor(Z)).abs().max().item() < 0.03
    assert torch.allclose(inv.R_inf.cpu(), torch.full((2,), 0.5, dtype=torch.float64), atol=0.05)
    rp = inv.predict_Rp()  # R1 + Rd = 1.7
    assert torch.allclose(rp.cpu(), torch.full((2,), 1.7, dtype=torch.float64), atol=0.15)
    with pytest.raises(NotImplementedError):
        inv.fit(freq, torch.tensor(Z), mode='optimize', nonneg=True, outliers=True)


def test_unsupported_options_are_loud():
    from bayes_drt_b200 import Inverter
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    inv = Inverter()
    for kw in (dict(part='real'), dict(fitY=True), dict(model_str='Series_StanModel.pkl'), dict(add_stan_data={'a': 1})):
        with pytest.raises(NotImplementedError):
            inv.fit(freq, Z, **kw)
    with pytest.raises(ValueError):
        inv.fit(freq, Z, mode='laplace')
    with pytest.raises(ValueError):
        inv.fit(freq[:-1], Z)
    with pytest.raises(ValueError):
        Inverter(basis='Zic')
    two_par = {'kernel': 'DDT', 'dist_type': 'parallel', 'symmetry': 'planar', 'bc': 'transmissive'}
    with pytest.raises(NotImplementedError):  # MultiDist placeholder of the reference (model file not shipped)
        Inverter(distributions={'a': dict(two_par), 'b': dict(two_par), 'c': dict(two_par)}).fit(freq, Z)
    with pytest.raises(NotImplementedError):  # Parallel_outliers is inconsistent as shipped
        Inverter(distributions={'a': dict(two_par)}).fit(freq, Z, outliers=True)
    with pytest.raises(ValueError):
        inv.coef_percentile('DRT', 50)  # no bayes fit yet


def test_init_from_ridge_and_auto_outliers():
    """The reference's recommended flow (inversion.py:1154-1187, :1616-1682): ridge solution -> Stan initial values;
    outliers='auto' switches spectra with IQR-flagged residuals to the outlier-robust model."""
    from bayes_drt_b200 import Inverter
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    Zo = Z.copy()
    Zo[40] += 0.15 * (1 + 1j)  # two gross outliers
    Zo[41] -= 0.12j
    Zb = torch.tensor(np.stack([Z, Zo]))
    rnd = Inverter()
    rnd.fit(freq, Zb, mode='optimize')
    inv = Inverter()
    inv.fit(freq, Zb, mode='optimize', init_from_ridge=True, outliers='auto')
    assert inv._outlier_model.tolist() == [False, True]
    assert inv.stan_model_name == 'Series_outliers_StanModel.pkl'
    so = inv.error_fit['sigma_out']
    assert torch.isnan(so[0]).all() and torch.isfinite(so[1]).all()
    # the flagged points carry most of the outlier variance
    assert set(torch.topk(so[1], 2).indices.tolist()) == {40, 41}
    # ridge initialisation only fixes x / R_inf / inductance (the hyper-parameters still start at random, as in Stan's
    # partial init), so it buys consistency rather than speed: same DRT on the clean spectrum
    c0, c1 = inv.distribution_fits['DRT']['coef'][0], rnd.distribution_fits['DRT']['coef'][0]
    assert (c0 - c1).abs().max().item() < 0.03 * c1.abs().max().item()
    # the robust fit of the contaminated spectrum stays close to the clean fit; the plain model is visibly distorted
    c_rob, c_plain = inv.distribution_fits['DRT']['coef'][1], rnd.distribution_fits['DRT']['coef'][1]
    assert (c_rob - c0).abs().max().item() < (c_plain - c0).abs().max().item()
    assert torch.isfinite(inv.R_inf).all() and abs(inv.R_inf[1].item() - 1.0) < 0.03
    # single-spectrum call warns like the reference
    one = Inverter()
    with pytest.warns(UserWarning, match='likely outliers'):
        one.fit(freq, Zo, mode='optimize', outliers='auto')
    assert one.stan_model_name == 'Series_outliers_StanModel.pkl' and one.error_fit['sigma_out'].shape == (81,)
    # nonneg + ridge init (exact zeros of the QP are floored before the log transform)
    pos = Inverter()
    pos.fit(freq, Z, mode='optimize', nonneg=True, init_from_ridge=True)
    assert np.isfinite(pos.distribution_fits['DRT']['coef']).all() and abs(pos.R_inf - 1.0) < 0.02


def test_save_and_load_fit_data(tmp_path):
    """inversion.py:3980-4064: the attribute dict survives a pickle round trip (numpy inside, no GPU needed to read it)
    and a fresh Inverter answers the post-fit queries from it."""
    import pickle
    from bayes_drt_b200 import Inverter
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    Zb = torch.tensor(np.stack([Z, load_spectrum('2ZARC_uniform_0.25')[1]]))
    inv = Inverter()
    inv.fit(freq, Zb, mode='sample', chains=2, warmup=60, samples=40)
    core = inv.save_fit_data(which='core')
    assert set(core) >= {'distributions', 'distribution_fits', 'f_train', 'Z_train', '_Z_scale', 'fit_type', 'R_inf',
                         'inductance', 'stan_model_name', '_sample_result', 'error_fit'}
    assert 'distribution_matrices' not in core and isinstance(core['distribution_fits']['DRT']['coef'], np.ndarray)
    fn = str(tmp_path / 'fit.pkl')
    inv.save_fit_data(fn, which='all')
    with open(fn, 'rb') as f:
        raw = pickle.load(f)
    assert 'distribution_matrices' in raw and isinstance(raw['Z_train'], np.ndarray)
    new = Inverter().load_fit_data(fn)
    assert new.fit_type == 'bayes' and new.stan_model_name == inv.stan_model_name
    tau = np.logspace(-6, 1, 50)
    for kw in (dict(), dict(percentile=97.5)):
        assert torch.equal(new.predict_distribution('DRT', eval_tau=tau, **kw), inv.predict_distribution('DRT', eval_tau=tau, **kw))
    assert torch.equal(new.predict_Z(freq), inv.predict_Z(freq))
    assert torch.equal(new.predict_Rp(percentile=50), inv.predict_Rp(percentile=50))
    # single-spectrum MAP fit, core only, through a dict
    one = Inverter()
    one.fit(freq, Z, mode='optimize', max_iter=500)
    again = Inverter().load_fit_data(one.save_fit_data(which='core'))
    assert np.array_equal(again.predict_distribution('DRT', eval_tau=tau), one.predict_distribution('DRT', eval_tau=tau))
    assert again.predict_Rp() == one.predict_Rp()
    with pytest.raises(ValueError):
        Inverter().save_fit_data()


def _zarc_batch(deltas, seed=3):
    """ZARC spectra on per-spectrum grids freq_b = 10**(6 - delta_b - arange(81)/10) (SURVEY 8d config 5 grids)."""
    rng = np.random.RandomState(seed)
    freq = 10.0 ** (6 - np.asarray(deltas)[:, None] - np.arange(81)[None, :] / 10)
    B = len(deltas)
    R0, R1 = rng.uniform(0.5, 2, B), rng.uniform(0.5, 2, B)
    tau0, n = 10.0 ** rng.uniform(-4, 0, B), rng.uniform(0.6, 1.0, B)
    Z = R0[:, None] + R1[:, None] / (1 + (2j * np.pi * freq * tau0[:, None]) ** n[:, None])
    Z = Z + 0.0025 * R1[:, None] * (rng.randn(B, 81) + 1j * rng.randn(B, 81))
    return freq, Z


def test_fit_per_spectrum_frequency_grids():
    """frequencies [B, Nf]: every spectrum on its own grid (matrices built per spectrum on the GPU) gives what the
    spectrum gives when fitted alone on its grid through the shared-grid path."""
    from bayes_drt_b200 import Inverter
    freq, Z = _zarc_batch([0.0, 0.13, 0.37, 0.5, 0.82])
    inv = Inverter()
    inv.fit(freq, Z, mode='optimize', polish=True)
    coef = inv.distribution_fits['DRT']['coef']
    assert tuple(coef.shape) == (5, 101) and inv.f_train.shape == (5, 81)
    assert tuple(np.shape(inv.distributions['DRT']['tau'])) == (5, 101)
    # (the random start of global spectrum 0 ends in a poor local mode that the Newton polish cannot leave -- on both
    # paths alike, which is what this test compares)
    assert int((inv._opt_result['gnorm'] < 1e-6).sum()) >= 4
    Zp = inv.predict_Z(inv.f_train)
    assert tuple(Zp.shape) == (5, 81) and float((Zp.cpu() - torch.tensor(Z))[1:].abs().max()) < 0.03
    gam = inv.predict_distribution()
    assert tuple(gam.shape) == (5, 101)
    rp = inv.predict_Rp()
    s_re, s_im = inv.predict_sigma(inv.f_train)
    assert tuple(s_re.shape) == (5, 81) and (s_re > 0).all() and (s_im > 0).all()
    assert inv.check_outliers(threshold=3.5).shape[1] == 2
    for b in (0, 1, 4):
        one = Inverter()
        one.fit(freq[b], Z[b], mode='optimize', polish=True, spectrum_offset=b)  # same global index -> same init
        c1 = one.distribution_fits['DRT']['coef']
        assert np.allclose(one.distributions['DRT']['tau'], inv.distributions['DRT']['tau'][b], rtol=1e-14)
        assert np.max(np.abs(c1 - coef[b].cpu().numpy())) <= 1e-5 * np.abs(c1).max(), b
        assert abs(one.R_inf - float(inv.R_inf[b])) <= 1e-5 * abs(one.R_inf)
        assert abs(one.predict_Rp() - float(rp[b])) <= 1e-5 * abs(one.predict_Rp())
        assert np.allclose(one.predict_distribution(), gam[b].cpu().numpy(), rtol=0, atol=1e-5 * np.abs(c1).max())
    # a shared basis for shifted grids, HMC started at the MAP estimates, percentiles
    inv2 = Inverter(basis_freq=np.logspace(7, -3, 101))
    u0 = inv._opt_result['u'][1:3, None, :].expand(-1, 2, -1).contiguous()
    inv2.fit(freq[1:3], Z[1:3], mode='sample', warmup=60, samples=40, chains=2, init=u0)
    assert inv2.fit_type == 'bayes' and tuple(inv2.distribution_fits['DRT']['coef'].shape) == (2, 101)
    lo, hi = inv2.coef_percentile('DRT', 2.5), inv2.coef_percentile('DRT', 97.5)
    assert (lo <= hi).all()
    Zq = inv2.predict_Z(inv2.f_train, percentile=50)
    assert tuple(Zq.shape) == (2, 81) and float((Zq.cpu() - torch.tensor(Z[1:3])).abs().max()) < 0.05
    # ridge-initialised MAP and automatic outlier detection on per-spectrum grids (inversion.py:1154-1187 row by row):
    # the batch splits by Stan program (one row gets the outlier model) and every row equals the same flow run on that
    # row alone (same global spectrum index, hence the same random part of the start)
    import warnings
    inv3 = Inverter()
    Zo = Z[:3].copy()
    Zo[1, 30] += 0.15 + 0.15j
    inv3.fit(freq[:3], Zo, init_from_ridge=True, outliers='auto', polish=True)
    assert inv3._outlier_model.cpu().tolist() == [False, True, False]
    assert tuple(np.shape(inv3.distributions['DRT']['tau'])) == (3, 101)  # the whole batch's grids are back in place
    for b in range(3):
        one = Inverter()
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            one.fit(freq[b:b + 1], Zo[b:b + 1], init_from_ridge=True, outliers='auto', polish=True, spectrum_offset=b)
        c1, cb = one.distribution_fits['DRT']['coef'][0], inv3.distribution_fits['DRT']['coef'][b]
        assert bool(one._outlier_model[0]) == (b == 1)
        # both runs are taken to the exact optimum (polish): an L-BFGS end point alone amplifies last-bit differences of
        # the inputs (torch's batched std() of |Z| depends on the batch shape) to the percent level
        # (the outlier model has 2 Nf more parameters, most of them on their bound: its polish stops at max|grad| < 1e-7
        # a few 1e-3 of the peak apart)
        tol = 1e-2 if b == 1 else 1e-4
        assert float((cb - c1).abs().max()) <= tol * float(c1.abs().max()), b
        assert float((inv3.R_inf[b] - one.R_inf[0]).abs()) <= tol * float(one.R_inf[0])
    assert 30 in torch.nonzero(inv3.error_fit['sigma_out'][1] > 5 * inv3.error_fit['sigma_out'][1].median())[:, 0].tolist()
    with pytest.raises(ValueError):
        inv2.fit(freq[:2], Z[0])


def test_check_outliers_ridge_branch():
    """check_outliers without a Stan fit: ridge fit (preset 'Huang') + inter-quartile rule (inversion.py:3351-3367)."""
    from bayes_drt_b200 import Inverter
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    Zo = Z.copy()
    Zo[30] += 0.2 + 0.2j
    inv = Inverter()
    idx = inv.check_outliers(freq, Zo, threshold=4, use_existing_fit=True)
    assert inv.fit_type == 'ridge' and 30 in idx and len(idx) <= 3
    # an existing ridge fit of the same data is reused; batch form returns (spectrum, frequency) pairs
    assert np.array_equal(inv.check_outliers(freq, Zo, threshold=4, use_existing_fit=True), idx)
    pairs = inv.check_outliers(freq, np.stack([Z, Zo]), threshold=4, use_existing_fit=False)
    assert pairs.shape[1] == 2 and [1, 30] in pairs.cpu().tolist() and [0, 30] not in pairs.cpu().tolist()
    # after a MAP fit of the same data the fitted error model is used; a new data set falls back to ridge
    inv.fit(freq, Zo, mode='optimize', outliers=True)
    assert inv.fit_type == 'map'
    zs = inv.check_outliers(freq, Zo, threshold=3.5, use_existing_fit=True)
    assert inv.fit_type == 'map' and isinstance(zs, np.ndarray)
    inv.check_outliers(freq, Z, threshold=4, use_existing_fit=True)
    assert inv.fit_type == 'ridge'


def test_predict_sigma_off_grid_score_and_weight_forms():
    """predict_sigma away from the training grid (inversion.py:3104-3137), score (:3141-3160), array / scalar /
    'prop_adj' weights of ridge_fit (_format_weights :2338-2395)."""
    from bayes_drt_b200 import Inverter
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    inv = Inverter()
    inv.fit(freq, Z, mode='optimize')
    s_re, s_im = inv.predict_sigma(freq)
    # the same error model rebuilt from its parameters at (almost) the training frequencies
    t_re, t_im = inv.predict_sigma(freq * (1 + 1e-6))
    assert t_re.shape == (81,) and np.allclose(t_re, s_re, rtol=1e-4) and np.allclose(t_im, s_im, rtol=1e-4)
    f2 = np.logspace(5.5, -1.5, 33)
    u_re, u_im = inv.predict_sigma(f2)
    assert u_re.shape == (33,) and (u_re > 0).all() and (u_im > 0).all()
    Zp = inv.predict_Z(freq)
    chi = inv.score(freq, Z)
    ref = np.sum(np.concatenate(((Zp - Z).real, (Zp - Z).imag)) ** 2) / 81
    assert isinstance(chi, float) and abs(chi - ref) <= 1e-12 * ref
    w = 1 / np.abs(Z)
    chi_w = inv.score(freq, Z, weights='modulus', part='imag')
    assert abs(chi_w - np.sum(((Zp - Z).imag * w) ** 2) / 81) <= 1e-12 * chi_w
    assert 0.999 < inv.score(freq, Z, metric='r2') <= 1.0
    with pytest.raises(ValueError):
        inv.score(freq, Z, metric='mse')
    # HMC fit: percentile version off the grid, batch shapes
    Zb = np.stack([Z, load_spectrum('ZARC-RL_uniform_0.25')[1]])
    hm = Inverter()
    hm.fit(freq, Zb, mode='sample', warmup=60, samples=40, chains=2, init_from_ridge=True)
    lo_re, lo_im = hm.predict_sigma(f2, percentile=10)
    hi_re, hi_im = hm.predict_sigma(f2, percentile=90)
    assert tuple(lo_re.shape) == (2, 33) and bool((lo_re > 0).all()) and bool((hi_im > 0).all())
    assert tuple(hm.score(freq, Zb).shape) == (2,)
    # every draw's impedance; its per-frequency median is what predict_Z(percentile=50) reports
    Zd = hm.predict_Z_distribution(f2)
    assert tuple(Zd.shape) == (2, 80, 33)
    med = hm.predict_Z(f2, percentile=50)
    assert torch.allclose(torch.quantile(Zd.real, 0.5, dim=1), med.real, rtol=0, atol=1e-12)
    assert torch.allclose(torch.quantile(Zd.imag, 0.5, dim=1), med.imag, rtol=0, atol=1e-12)
    # accessors of the reference (inversion.py:133, :4069-4110)
    assert hm.get_distributions() is hm.distributions and hm.get_basis() == 'gaussian' and hm.get_fit_inductance()
    hm.set_epsilon(3.0, override_distributions=True)
    assert hm.get_epsilon() == 3.0 and hm.distributions['DRT']['epsilon'] == 3.0 and hm.get_basis_freq() is None
    lam = Inverter().ridge_ReImCV(freq, Z, lambdas=np.logspace(-5, -1, 5))
    assert isinstance(lam, float) and 1e-5 <= lam <= 1e-1
    # ridge weights: an array equal to the 'modulus' weights of the scaled data reproduces weights='modulus'
    r1 = Inverter()
    r1.ridge_fit(freq, Z, weights='modulus')
    c1 = r1.distribution_fits['DRT']['coef'].copy()
    r1.ridge_fit(freq, Z, weights=1 / np.abs(Z / float(r1._Z_scale[0])))
    assert np.max(np.abs(r1.distribution_fits['DRT']['coef'] - c1)) <= 1e-9 * np.abs(c1).max()
    r1.ridge_fit(freq, Z, weights=(1 + 1j) / np.abs(Z / float(r1._Z_scale[0])))
    assert np.max(np.abs(r1.distribution_fits['DRT']['coef'] - c1)) <= 1e-9 * np.abs(c1).max()
    for wt in (2.0, 'prop_adj', 1 + 2j):
        r1.ridge_fit(freq, Z, weights=wt)
        assert np.isfinite(r1.distribution_fits['DRT']['coef']).all() and abs(r1.predict_Rp() - 1.0) < 0.1
    with pytest.raises(ValueError):
        r1.ridge_fit(freq, Z, weights=np.ones(5))


def test_fit_init_shapes_failures_and_diagnostics():
    """a [B, D] init with several chains starts every chain there; a bad explicit init is an error
    (Stan: "Initialization failed"); chains * samples beyond what the percentile kernel holds is refused before the
    sampler runs; HMC fits carry split R-hat and bulk ESS of every coefficient, R_inf and the inductance."""
    from bayes_drt_b200 import Inverter, synth
    freq, Z, _ = synth.make_spectra(3, seed=8)
    _, bf = synth.bench_grid()
    inv = Inverter(basis_freq=bf.numpy())
    D = 2 * 100 + 9
    u = torch.zeros(3, D, dtype=torch.float64)
    inv.fit(freq, Z, mode='sample', chains=3, warmup=30, samples=20, init=u, check_outliers=False)
    st = inv._sample_stats
    assert tuple(st['rhat'].shape) == (3, 102) and tuple(st['ess_bulk'].shape) == (3, 102)
    assert torch.isfinite(st['rhat']).all() and (st['ess_bulk'] > 0).all()
    with pytest.raises(ValueError, match='init must be'):
        inv.fit(freq, Z, mode='sample', chains=3, warmup=30, samples=20, init=torch.zeros(3, 2, D, dtype=torch.float64))
    bad = u.clone()
    bad[1, 5] = float('nan')
    with pytest.raises(RuntimeError, match='Initialization failed'):
        inv.fit(freq, Z, mode='optimize', init=bad)
    with pytest.raises(ValueError, match='merged draws'):
        inv.fit(freq, Z, mode='sample', chains=4, warmup=10, samples=5000)
    # the public preparation hook hands out the same problem / starts that fit() uses
    prob, u0 = inv.prepare(freq, Z, mode='optimize')
    r = prob.map_lbfgs(u0, max_iter=50)
    inv.fit(freq, Z, mode='optimize', max_iter=50, check_outliers=False)
    assert torch.equal(r['u'], inv._opt_result['u'])


def test_hmc_in_pieces_equals_one_launch():
    """keep_draws=False runs a large batch through the sampler in bounded pieces (memory of the draws): same posterior
    means, diagnostics and sampler statistics as one launch, bit for bit (streams keyed by the global spectrum index)."""
    from bayes_drt_b200 import Inverter, synth
    freq, Z, _ = synth.make_spectra(7, seed=9)
    _, bf = synth.bench_grid()
    kw = dict(mode='sample', chains=2, warmup=30, samples=20, check_outliers=False, spectrum_offset=40)
    one = Inverter(basis_freq=bf.numpy())
    one.fit(freq, Z, keep_draws=True, **kw)
    pcs = Inverter(basis_freq=bf.numpy())
    pcs._hmc_piece = 3
    pcs.fit(freq, Z, keep_draws=False, **kw)
    assert pcs._sample_result is None
    assert torch.equal(one.distribution_fits['DRT']['coef'], pcs.distribution_fits['DRT']['coef'])
    assert torch.equal(one.R_inf, pcs.R_inf)
    for k in ('stepsize', 'n_leapfrog', 'rhat', 'ess_bulk'):
        assert torch.equal(one._sample_stats[k], pcs._sample_stats[k]), k
