"""CPU: the shipped Inverter end to end with the device library replaced by the oracle-backed test double
(tests/cpu_seam.py): a guard for the Python host code -- argument handling, reference shapes, model choice, post-fit
queries, persistence -- that runs without a GPU.  Numerical parity of the CUDA path is the GPU tests' job."""
import numpy as np
import pytest
import torch

import cpu_seam
from helpers import load_spectrum


@pytest.fixture
def inverter(monkeypatch):
    return cpu_seam.install(monkeypatch)


def test_map_fit_single_spectrum_reference_shapes_and_queries(inverter, tmp_path):
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    inv = inverter.Inverter()
    inv.fit(freq, Z, mode='optimize')
    assert inv.fit_type == 'map' and inv.stan_model_name == 'Series_StanModel.pkl'
    coef = inv.distribution_fits['DRT']['coef']
    assert isinstance(coef, np.ndarray) and coef.shape == (101,) and isinstance(inv.R_inf, float)
    assert abs(inv.R_inf - 1.0) < 0.02 and abs(inv.predict_Rp() - 1.0) < 0.05  # truth: R_inf = R_p = 1
    assert inv.error_fit['sigma_tot'].shape == (162,)
    Zp = inv.predict_Z(freq)
    assert Zp.shape == (81,) and np.max(np.abs(Zp - Z)) < 0.02
    assert inv.predict_distribution('DRT', eval_tau=np.logspace(-6, 1, 50)).shape == (50,)
    s_re, s_im = inv.predict_sigma(freq)
    assert s_re.shape == (81,) and (s_re > 0).all() and (s_im > 0).all()
    assert 0 < inv.score(freq, Z) < 1e-3 and inv.score(freq, Z, metric='r2') > 0.999
    assert isinstance(inv.check_outliers(threshold=3.5), np.ndarray)
    # persistence round trip (inversion.py:3980-4064)
    fn = str(tmp_path / 'fit.pkl')
    inv.save_fit_data(fn)
    other = inverter.Inverter().load_fit_data(fn)
    assert np.allclose(other.predict_Z(freq), Zp) and other.fit_type == 'map'


def test_map_fit_batch_outlier_model_and_loud_errors(inverter):
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    Zo = Z.copy()
    Zo[40] += 0.15 * (1 + 1j)
    inv = inverter.Inverter()
    inv.fit(freq, np.stack([Z, Zo]), mode='optimize', outliers=True, max_iter=3000)
    assert inv.stan_model_name == 'Series_outliers_StanModel.pkl'
    assert tuple(inv.distribution_fits['DRT']['coef'].shape) == (2, 101) and tuple(inv.R_inf.shape) == (2,)
    so = inv.error_fit['sigma_out']
    assert tuple(so.shape) == (2, 81) and bool(torch.isfinite(so).all()) and bool((so > 0).all())
    for kw, exc in ((dict(part='real'), NotImplementedError), (dict(fitY=True), NotImplementedError),
                    (dict(mode='anneal'), ValueError)):
        with pytest.raises(exc):
            inv.fit(freq, Z, **kw)
    with pytest.raises(ValueError):
        inv.fit(freq[:-1], Z)


def test_sample_fit_percentiles_and_draw_queries(inverter):
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    inv = inverter.Inverter(basis_freq=np.logspace(5, -1, 31))  # small basis: the oracle's NUTS is plain numpy
    mp = inverter.Inverter(basis_freq=np.logspace(5, -1, 31))
    mp.fit(freq[::2], Z[::2], mode='optimize', max_iter=4000)
    u0 = mp._opt_result['u'][:, None, :].expand(-1, 2, -1)
    inv.fit(freq[::2], Z[::2], mode='sample', chains=2, warmup=30, samples=20, init=u0)
    assert inv.fit_type == 'bayes' and inv.distribution_fits['DRT']['coef'].shape == (31,)
    lo, hi = inv.coef_percentile('DRT', 2.5), inv.coef_percentile('DRT', 97.5)
    assert lo.shape == (31,) and (lo <= hi).all()
    Zd = inv.predict_Z_distribution(freq[::2])
    assert Zd.shape == (40, 41)
    med = inv.predict_Z(freq[::2], percentile=50)
    assert np.allclose(np.percentile(Zd.real, 50, axis=0), med.real, atol=1e-12)
    assert isinstance(inv.predict_Rp(percentile=50), float)
    s_lo, _ = inv.predict_sigma(freq[::2], percentile=10)
    s_hi, _ = inv.predict_sigma(freq[::2], percentile=90)
    assert (s_lo <= s_hi).all()


def test_sample_fit_in_pieces_equals_one_batch(inverter):
    """keep_draws=False runs a large HMC batch through the sampler in bounded pieces (prob.subset) and summarises each
    before the next: same posterior means, diagnostics and sampler statistics as the one-batch run (the random streams
    are keyed by the global spectrum index), and no draws are kept."""
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    f, Zb = freq[::4], np.stack([Z[::4], 1.1 * Z[::4], 0.9 * Z[::4]])
    bf = np.logspace(5, -1, 21)
    u0 = np.zeros((3, 2, 2 * 21 + 9))
    kw = dict(mode='sample', chains=2, warmup=15, samples=10, init=torch.tensor(u0), check_outliers=False)
    one = inverter.Inverter(basis_freq=bf)
    one.fit(f, Zb, **kw)
    pieces = inverter.Inverter(basis_freq=bf)
    pieces._hmc_piece = 2  # pieces [0:2] and [2:3]
    pieces.fit(f, Zb, keep_draws=False, **kw)
    assert np.allclose(pieces.distribution_fits['DRT']['coef'], one.distribution_fits['DRT']['coef'], rtol=1e-12)
    assert np.allclose(pieces.R_inf, one.R_inf, rtol=1e-12)
    for k in ('stepsize', 'n_leapfrog', 'rhat', 'ess_bulk'):
        assert np.allclose(np.asarray(pieces._sample_stats[k]), np.asarray(one._sample_stats[k]), equal_nan=True), k
    assert pieces._sample_result is None and one._sample_result is not None
    with pytest.raises(Exception):
        pieces.coef_percentile('DRT', 50)


def test_multi_distribution_fits_shapes_and_scaling(inverter):
    """Series-Parallel / Series-2Parallel / Parallel through the host code: coefficient blocks per distribution, series
    coefficients scaled up and parallel ones scaled down by Z_scale (inversion.py:2445-2450)."""
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    bf = np.logspace(5, -1, 21)
    tp = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'transmissive', 'dist_type': 'parallel', 'basis_freq': bf}
    bp = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'blocking', 'dist_type': 'parallel', 'basis_freq': bf}
    drt = {'kernel': 'DRT', 'basis_freq': bf}
    f, z = freq[::2], Z[::2]
    sp = inverter.Inverter(distributions={'DRT': dict(drt), 'TP-DDT': dict(tp, x_scale=0.8)})
    sp.fit(f, z, mode='optimize', nonneg=True, max_iter=300)
    assert sp.stan_model_name == 'Series-Parallel_pos_StanModel.pkl'
    assert sp.distribution_fits['DRT']['coef'].shape == (21,) and sp.distribution_fits['TP-DDT']['coef'].shape == (21,)
    assert sp.predict_Z(f).shape == (41,) and sp.predict_Z(f, distributions='DRT', include_offsets=False).shape == (41,)
    s2 = inverter.Inverter(distributions={'DRT': dict(drt), 'TP-DDT': dict(tp, x_scale=0.8), 'BP-DDT': dict(bp)})
    s2.fit(f, np.stack([z, z]), mode='optimize', nonneg=True, max_iter=200)
    assert s2.stan_model_name == 'Series-2Parallel_pos_StanModel.pkl'
    assert s2.distributions['BP-DDT']['order'] == 1 and s2.distributions['TP-DDT']['order'] == 2  # by name
    assert all(tuple(s2.distribution_fits[k]['coef'].shape) == (2, 21) for k in ('DRT', 'TP-DDT', 'BP-DDT'))
    par = inverter.Inverter(distributions={'TP-DDT': dict(tp)})
    par.fit(f, z, mode='optimize', max_iter=300)
    assert par.stan_model_name == 'Parallel_StanModel.pkl' and float(par._Z_scale[0]) > 10  # admittance scaling
    assert par.distribution_fits['TP-DDT']['coef'].shape == (21,) and (par.distribution_fits['TP-DDT']['coef'] > 0).all()
    with pytest.raises(NotImplementedError):
        sp.fit(f, z, outliers=True)


def test_hooks_bench_py_relies_on(inverter):
    """bench.py assembles its resident-input leg from Inverter's own preprocessing (_to_batch, _scale_Z, _grid) and reads
    _opt_result / _sample_result / _sample_stats: keep those hooks working."""
    from bayes_drt_b200 import synth
    freq, Z, _ = synth.make_spectra(3, seed=20240601)
    _, bf = synth.bench_grid()
    inv = inverter.Inverter(basis_freq=bf.numpy(), device=None)
    fs, Zb = inv._to_batch(freq, Z)
    Zs = inv._scale_Z(Zb, True)
    tau, eps, m = inv._grid(fs, 'DRT')
    assert tuple(Zs.shape) == (3, 70) and len(tau) == 100 and abs(eps - 4.342944819) < 1e-8
    assert {'A_re', 'A_im', 'L0', 'L1', 'L2'} <= set(m) and tuple(m['A_re'].shape) == (70, 100)
    inv.fit(freq, Z[:2], mode='optimize', max_iter=30)
    assert {'u', 'lp', 'iters', 'n_eval', 'status'} <= set(inv._opt_result)
    hm = inverter.Inverter(basis_freq=np.logspace(5, -1, 13), device=None)
    hm.fit(freq, Z[:1], mode='sample', chains=2, warmup=6, samples=4, check_outliers=False, spectrum_offset=5)
    assert tuple(hm._sample_result['x'].shape) == (1, 8, 13)
    assert {'stepsize', 'n_leapfrog', 'n_divergent', 'n_maxdepth', 'accept'} <= set(hm._sample_stats)
