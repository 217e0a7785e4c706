"""CPU: what the shipped host code of bayes_drt_b200.Inverter.fit hands to the CUDA solvers, against what the
reference's own Inverter.fit hands to Stan (tests/golden/stan_data.npz, scripts/make_golden_stan_data.py).

The device library is replaced here by a recorder: kernel matrices come from the oracle (their CUDA builders are
tested against the same reference matrices on the GPU), capi.SeriesProblem captures its arguments and aborts the fit.
Everything between the user's call and that point -- sorting, scaling, default grids, model choice, constants of each
mode, stacking, the parallel families -- is the shipped code, running on CPU tensors."""
import os

import numpy as np
import pytest
import torch

from oracle import matrices as om

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'stan_data.npz'))
FREQ, Z = G['freq'], G['Z']
BF = np.logspace(6, -2, 81)
TP = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'transmissive', 'dist_type': 'parallel', 'basis_freq': BF}
BP = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'blocking', 'dist_type': 'parallel', 'basis_freq': BF}
DRT = {'kernel': 'DRT', 'basis_freq': BF}
CASES = {
    'series_opt': (dict(), dict(mode='optimize')),
    'series_sample': (dict(), dict(mode='sample')),
    'series_pos_opt': (dict(), dict(mode='optimize', nonneg=True)),
    'series_out_opt': (dict(), dict(mode='optimize', outliers=True)),
    'series_pos_out_sample': (dict(), dict(mode='sample', nonneg=True, outliers=True, outlier_lambda=5, sigma_min=0.001,
                                           inductance_scale=2)),
    'series_basis_eq_freq': (dict(basis_freq=FREQ), dict(mode='optimize')),
    'series_noscale': (dict(), dict(mode='optimize', scale_Z=False)),
    'sp_opt': (dict(distributions={'DRT': dict(DRT), 'TP-DDT': dict(TP, x_scale=0.8)}), dict(mode='optimize', nonneg=True)),
    'sp_sample': (dict(distributions={'DRT': dict(DRT), 'TP-DDT': dict(TP, x_scale=0.8)}), dict(mode='sample', nonneg=True)),
    'parallel_tp_opt': (dict(distributions={'TP-DDT': dict(TP)}), dict(mode='optimize')),
    'parallel_bp_sample': (dict(distributions={'BP-DDT': dict(BP)}), dict(mode='sample')),
    's2p_opt': (dict(distributions={'DRT': dict(DRT), 'TP-DDT': dict(TP, x_scale=0.8), 'BP-DDT': dict(BP)}),
                dict(mode='optimize', nonneg=True)),
    's2p_sample': (dict(distributions={'DRT': dict(DRT), 'TP-DDT': dict(TP, x_scale=0.8), 'BP-DDT': dict(BP)}),
                   dict(mode='sample', nonneg=True)),
}


class Abort(Exception):
    pass


class Recorder:
    """stands in for capi.SeriesProblem: keeps the constructor arguments, stops the fit at the first solver call"""
    last = None

    def __init__(self, A, Z, freq, L, **kw):
        self.A, self.Z, self.freq, self.L, self.kw = A, Z, freq, L, kw
        self.B, self.Nf, self.K = Z.shape[0], Z.shape[1] // 2, A.shape[-1]
        n_par = [kw[k].shape[-1] for k in ('Ap', 'Ap2') if kw.get(k) is not None]
        self.D = 2 * (self.K + sum(n_par)) + 6 + 3 * (1 + len(n_par)) + (2 * self.Nf if kw.get('outliers') else 0)
        Recorder.last = self

    def logpost_grad(self, u, spec=None, jacobian=False):
        """the initial-point check of fit(): every point is accepted here"""
        return torch.zeros(u.shape[0], dtype=torch.float64), torch.zeros_like(u)

    def map_lbfgs(self, *a, **k):
        self.call = ('optimizing', k)
        raise Abort

    def nuts(self, *a, **k):
        self.call = ('sampling', k)
        raise Abort


@pytest.fixture
def host(monkeypatch):
    from bayes_drt_b200 import inverter

    def build_A(freq, tau, eps, kernel='DRT', dist_type='series', symmetry='planar', bc='transmissive', ct=False,
                k_ct=None, device=None):
        kw = dict(tau=np.asarray(tau, dtype=np.float64), epsilon=float(eps), kernel=kernel, dist_type=dist_type,
                  symmetry=symmetry, bc=bc, ct=ct, k_ct=k_ct)
        f = np.asarray(freq, dtype=np.float64)
        if f.ndim == 2:  # one grid per row (capi.build_A: freq [G, Nf], tau [K] or [G, K])
            t = kw.pop('tau')
            rows = [(om.construct_A(f[g], 'real', tau=t[g] if t.ndim == 2 else t, **kw),
                     om.construct_A(f[g], 'imag', tau=t[g] if t.ndim == 2 else t, **kw)) for g in range(f.shape[0])]
            return torch.tensor(np.stack([r[0] for r in rows])), torch.tensor(np.stack([r[1] for r in rows]))
        return torch.tensor(om.construct_A(f, 'real', **kw)), torch.tensor(om.construct_A(f, 'imag', **kw))

    def build_L(freq, tau, eps, order, device=None):
        return torch.tensor(om.construct_L(np.asarray(freq, dtype=np.float64), tau=np.asarray(tau, dtype=np.float64),
                                           epsilon=float(eps), order=order))
    monkeypatch.setattr(inverter, 'context', lambda device=None: type('Ctx', (), {'device': torch.device('cpu')})())
    monkeypatch.setattr(inverter.capi, 'build_A', build_A)
    monkeypatch.setattr(inverter.capi, 'build_L', build_L)
    monkeypatch.setattr(inverter.capi, 'SeriesProblem', Recorder)
    return inverter


def _proj(case, key, M, rtol=1e-11):
    M = np.asarray(M, dtype=np.float64)
    r = np.random.RandomState(M.shape[0] * 1000 + M.shape[1])
    v, u = r.standard_normal(M.shape[1]), r.standard_normal(M.shape[0])
    a, b = G[f'{case}/{key}@v'], G[f'{case}/u@{key}']
    assert np.max(np.abs(M @ v - a)) <= rtol * np.max(np.abs(a)), (case, key)
    assert np.max(np.abs(u @ M - b)) <= rtol * np.max(np.abs(b)), (case, key)


@pytest.mark.parametrize('case', sorted(CASES))
def test_fit_hands_the_solver_what_the_reference_hands_stan(case, host):
    ikw, fkw = CASES[case]
    inv = host.Inverter(**ikw)
    with pytest.raises(Abort):
        inv.fit(FREQ[::-1].copy(), Z[::-1].copy(), **fkw)  # ascending input: the sort is part of what is checked
    p, kw = Recorder.last, Recorder.last.kw
    assert p.call[0] == ('optimizing' if fkw['mode'] == 'optimize' else 'sampling')
    # the Stan program the reference selects (Inverter._get_stan_model) <-> the model flags of the C ABI
    ref_model = str(G[f'{case}/model'])[:-len('_StanModel.pkl')]
    family = 'Series-2Parallel' if kw.get('Ap2') is not None else ('Series-Parallel' if kw.get('Ap') is not None else
                                                                   ('Parallel' if kw.get('parallel') else 'Series'))
    pos = kw['nonneg'] and family != 'Parallel'
    assert ref_model == family + ('_pos' if pos else '') + ('_outliers' if kw.get('outliers') else '')
    assert float(inv._Z_scale[0]) == pytest.approx(float(G[f'{case}/Z_scale']), rel=1e-13)
    assert np.array_equal(p.freq.numpy(), G[f'{case}/freq'])
    assert np.max(np.abs(p.Z[0].numpy() - G[f'{case}/Z'])) <= 1e-13 * np.max(np.abs(G[f'{case}/Z']))
    names = {'Series': ('A', ['L0', 'L1', 'L2']), 'Parallel': ('A', ['L0', 'L1', 'L2']),
             'Series-Parallel': ('As', ['L0s', 'L1s', 'L2s']), 'Series-2Parallel': ('As', ['L0s', 'L1s', 'L2s'])}[family]
    _proj(case, names[0], p.A.numpy())
    for o, nm in enumerate(names[1]):
        _proj(case, nm, p.L[o].numpy())
    for k_ref, k in (('sigma_min', 'sigma_min'), ('ups_alpha', 'ups_alpha'), ('ups_beta', 'ups_beta'),
                     ('induc_scale', 'induc_scale')):
        assert float(G[f'{case}/{k_ref}']) == pytest.approx(float(kw[k]), rel=1e-14), k
    if kw.get('outliers'):
        for k in ('sigma_out_lambda', 'sigma_out_alpha', 'sigma_out_beta'):
            assert float(G[f'{case}/{k}']) == pytest.approx(float(kw[k]), rel=1e-14), k
    if family == 'Series-Parallel':
        _proj(case, 'Ap', kw['Ap'].numpy())
        for o in range(3):
            _proj(case, f'L{o}p', kw['Lp'][o].numpy())
        assert float(G[f'{case}/x_sum_invscale']) == kw['x_sum_invscale'] and float(G[f'{case}/xp_scale']) == kw['xp_scale']
    if family == 'Series-2Parallel':  # parallel distributions in the reference's by-name order
        _proj(case, 'Ap1', kw['Ap'].numpy())
        _proj(case, 'Ap2', kw['Ap2'].numpy())
        for o in range(3):
            _proj(case, f'L{o}p1', kw['Lp'][o].numpy())
            _proj(case, f'L{o}p2', kw['Lp2'][o].numpy())
        assert float(G[f'{case}/x_sum_invscale']) == kw['x_sum_invscale']
        assert float(G[f'{case}/xp1_scale']) == kw['xp_scale'] and float(G[f'{case}/xp2_scale']) == kw['xp2_scale']
    if fkw['mode'] == 'sample':  # inversion.py:1218-1221
        assert p.call[1]['chains'] == int(G[f'{case}/chains']) and p.call[1]['warmup'] == int(G[f'{case}/warmup'])
        assert p.call[1]['warmup'] + p.call[1]['samples'] == int(G[f'{case}/iter'])
        assert p.call[1]['seed'] == int(G[f'{case}/seed'])
    else:
        assert p.call[1]['max_iter'] == int(G[f'{case}/iter'])


RIDGE_CASES = [dict(), dict(preset='Huang'), dict(penalty='integral', lambda_0=1, hl_beta=5, weights='modulus'),
               dict(nonneg=False, weights='proportional'), dict(reg_ord=[0.2, 0.3, 0.5], L1_penalty=0.01),
               dict(penalty='cholesky', weights='Orazem'), dict(hl_fbeta=0.1, scale_Z=False)]


@pytest.mark.parametrize('kw', RIDGE_CASES)
def test_ridge_fit_hands_the_kernel_the_system_of_the_reference(kw, host, monkeypatch):
    """Host side of ridge_fit: the weighted, augmented system and the options that reach bdrt_ridge_fit, against the
    oracle's prep() -- whose results are pinned to the reference's own ridge_fit in ridge_reference.npz."""
    from bayes_drt_b200 import ridge
    from oracle import ridge as oridge
    rec = {}

    def ridge_fit(WA_re, WA_im, WZ_re, WZ_im, Pen, Lmat, **o):
        rec.update(WA_re=WA_re, WA_im=WA_im, WZ_re=WZ_re, WZ_im=WZ_im, Pen=Pen, Lmat=Lmat, o=o)
        raise Abort

    def build_M(freq, eps, order, toeplitz, device=None):
        return torch.tensor(om.construct_M(np.asarray(freq, dtype=np.float64), order=order, epsilon=float(eps)))
    monkeypatch.setattr(ridge.capi, 'ridge_fit', ridge_fit)
    monkeypatch.setattr(ridge.capi, 'build_M', build_M)
    inv = host.Inverter()
    with pytest.raises(Abort):
        inv.ridge_fit(FREQ[::-1].copy(), Z[::-1].copy(), **kw)
    k = dict(kw)
    if k.pop('preset', None) == 'Huang':
        k.update(penalty='integral', hl_beta=2.5, lambda_0=1e-2, weights='modulus')
    p = oridge.prep(FREQ, Z, penalty=k.get('penalty', 'discrete'), weights=k.get('weights'), scale_Z=k.get('scale_Z', True))
    close = lambda a, b: np.max(np.abs(np.asarray(a) - b)) <= 1e-12 * max(np.max(np.abs(b)), 1e-300)
    WA_re, WA_im = rec['WA_re'].numpy(), rec['WA_im'].numpy()
    assert close(WA_re[0] if WA_re.ndim == 3 else WA_re, p['WA_re']) and close(WA_im[0] if WA_im.ndim == 3 else WA_im, p['WA_im'])
    assert close(rec['WZ_re'][0].numpy(), p['WZ_re']) and close(rec['WZ_im'][0].numpy(), p['WZ_im'])
    assert close(rec['Pen'].numpy(), p['Pen'])
    if k.get('penalty', 'discrete') != 'integral':
        assert close(rec['Lmat'].numpy(), p['Lmat'])
    o = rec['o']
    assert o['penalty'] == ('integral' if k.get('penalty') == 'integral' else 'discrete')
    assert o['nonneg'] == k.get('nonneg', True) and o['max_iter'] == 20 and o['xtol'] == 1e-3
    assert o['hl_beta'] == k.get('hl_beta', 2.5) and o['lambda_0'] == k.get('lambda_0', 1e-2)
    frac = np.zeros(3)
    if isinstance(k.get('reg_ord', 2), int):
        frac[k.get('reg_ord', 2)] = 1
    else:
        frac[:] = k['reg_ord']
    assert np.array_equal(np.asarray(o['reg_ord'], dtype=np.float64), frac) and o['L1_penalty'] == k.get('L1_penalty', 0)
    assert o['epsilon'] == pytest.approx(p['epsilon'], rel=1e-14) and o['hl_fbeta'] == k.get('hl_fbeta')
    assert float(inv._Z_scale[0]) == pytest.approx(p['Z_scale'], rel=1e-14)


def test_check_outliers_ridge_branch_matches_the_reference(host, monkeypatch):
    """Inverter.check_outliers without a Stan fit (inversion.py:3313-3367): ridge fit with preset 'Huang', residuals
    relative to |Z|, inter-quartile rule.  The shipped host code runs on CPU with the device ridge solver replaced by
    the oracle's (pinned to the reference's ridge_fit); the flagged indices are the reference's own."""
    from bayes_drt_b200 import ridge
    from oracle import ridge as oridge
    R = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ridge_reference.npz'))
    Zo = R['outliers/Z']
    seen = {}

    def ridge_fit(inv, frequencies, Zb, **kw):
        seen.update(kw)
        f = np.asarray(frequencies, dtype=np.float64)
        Zn = np.asarray(Zb.cpu() if torch.is_tensor(Zb) else Zb).reshape(-1, len(f))
        rs = [oridge.ridge_fit(f, z, **kw) for z in Zn]
        inv.f_train, inv.Z_train = rs[0]['prep']['freq'], torch.tensor(Zn)
        inv._Z_scale = torch.tensor([r['Z_scale'] for r in rs], dtype=torch.float64)
        inv.distributions['DRT']['tau'], inv.distributions['DRT']['epsilon'] = rs[0]['prep']['tau'], rs[0]['prep']['epsilon']
        inv.distribution_fits = {'DRT': {'coef': torch.tensor(np.stack([r['coef'] for r in rs]))}}
        inv.R_inf = torch.tensor([r['R_inf'] for r in rs], dtype=torch.float64)
        inv.inductance = torch.tensor([r['inductance'] for r in rs], dtype=torch.float64)
        inv.fit_type, inv._recalc_mat = 'ridge', False
        return inv
    monkeypatch.setattr(ridge, 'ridge_fit', ridge_fit)
    for thr in (0.5, 1.5, 4):
        inv = host.Inverter()
        idx = inv.check_outliers(FREQ, Zo, threshold=thr, use_existing_fit=False)
        assert seen == {'preset': 'Huang'}
        assert np.array_equal(np.asarray(idx), R[f'outliers/idx_t{thr}']), thr
    # batch form: (spectrum, frequency) pairs; the clean spectrum contributes its own noise-level flags only
    inv = host.Inverter()
    pairs = inv.check_outliers(FREQ, np.stack([Z, Zo]), threshold=4, use_existing_fit=False)
    assert [int(j) for i, j in pairs.tolist() if i == 1] == list(R['outliers/idx_t4'])


def _oracle_ridge(monkeypatch):
    """replace the device ridge solver behind Inverter by the oracle's (pinned to the reference's ridge_fit)"""
    from bayes_drt_b200 import ridge
    from oracle import ridge as oridge

    def ridge_fit(inv, frequencies, Zb, **kw):
        f = np.asarray(frequencies, dtype=np.float64)
        Zn = np.asarray(Zb.cpu() if torch.is_tensor(Zb) else Zb).reshape(-1, len(f))
        kw = {k: v for k, v in kw.items() if k != 'hyper_lambda'}
        rs = [oridge.ridge_fit(f, z, **kw) for z in Zn]
        inv.f_train, inv.Z_train = rs[0]['prep']['freq'], torch.tensor(Zn)
        inv._Z_scale = torch.tensor([r['Z_scale'] for r in rs], dtype=torch.float64)
        inv.distributions['DRT']['tau'], inv.distributions['DRT']['epsilon'] = rs[0]['prep']['tau'], rs[0]['prep']['epsilon']
        inv.distribution_fits = {'DRT': {'coef': torch.tensor(np.stack([r['coef'] for r in rs]))}}
        inv.R_inf = torch.tensor([r['R_inf'] for r in rs], dtype=torch.float64)
        inv.inductance = torch.tensor([r['inductance'] for r in rs], dtype=torch.float64)
        inv.fit_type, inv._recalc_mat = 'ridge', False
        return inv
    monkeypatch.setattr(ridge, 'ridge_fit', ridge_fit)


class InitRecorder(Recorder):
    """also keeps the initial points handed to the solver"""
    def map_lbfgs(self, u0, **k):
        self.u0 = u0
        return super().map_lbfgs(u0, **k)

    def nuts(self, u0, **k):
        self.u0 = u0
        return super().nuts(u0, **k)


@pytest.mark.parametrize('case,mode,ridge_init', [('auto_contaminated', 'optimize', False),
                                                   ('auto_contaminated_ridge_init', 'optimize', True),
                                                   ('auto_clean_ridge_init', 'sample', True),
                                                   ('series_ridge_init', 'optimize', True),
                                                   ('auto_marginal_ridge_init', 'optimize', True),
                                                   ('true_marginal_ridge_init', 'optimize', True)])
def test_auto_outliers_and_ridge_initialisation_flow(case, mode, ridge_init, host, monkeypatch):
    """outliers='auto' (inversion.py:1171-1187) and init_from_ridge (:1154-1160, :1616-1682) in the shipped fit():
    the Stan program chosen for the spectrum, the data it gets, and the initial values taken from the ridge solution."""
    _oracle_ridge(monkeypatch)
    monkeypatch.setattr(host.capi, 'SeriesProblem', InitRecorder)
    Zin = G['Z_contaminated'] if 'contaminated' in case else (G['Z_marginal'] if 'marginal' in case else Z)
    inv = host.Inverter()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        with pytest.raises(Abort):
            inv.fit(FREQ, Zin, mode=mode, init_from_ridge=ridge_init,
                    outliers='auto' if case.startswith('auto') else case.startswith('true'))
    p = Recorder.last
    ref_model = str(G[f'{case}/model'])[:-len('_StanModel.pkl')]
    assert ref_model == 'Series' + ('_outliers' if p.kw.get('outliers') else '')
    assert np.max(np.abs(p.Z[0].numpy() - G[f'{case}/Z'])) <= 1e-13 * np.max(np.abs(G[f'{case}/Z']))
    K, Nf = p.K, p.Nf
    u0 = p.u0.reshape(-1, p.D)
    if not ridge_init:
        assert str(G[f'{case}/init']) == 'random' and bool((u0.abs() <= 2).all())  # Stan's U(-2, 2) start
        return
    for row in u0:  # every chain starts from the ridge solution; what it does not fix stays random
        row = row.numpy()
        x = G[f'{case}/init/x']
        assert np.max(np.abs(row[2:2 + K] - x)) <= 1e-9 * np.max(np.abs(x))
        assert row[0] == pytest.approx(np.log(float(G[f'{case}/init/Rinf_raw'])), rel=1e-9)
        assert row[1] == pytest.approx(np.log(float(G[f'{case}/init/induc_raw'])), rel=1e-9)
        if p.kw.get('outliers'):
            assert np.allclose(row[6 + K:6 + K + Nf], np.log(G[f'{case}/init/sigma_out_raw']), rtol=1e-12)
        rest = np.r_[row[2 + K:6 + K], row[6 + K + (2 * Nf if p.kw.get('outliers') else 0):]]
        assert np.all(np.abs(rest) <= 2)


@pytest.fixture
def host_ridge(host, monkeypatch):
    """the device entry points behind ridge_fit (bdrt_ridge_fit, bdrt_qp_bound, bdrt_build_M) served by the oracle"""
    from bayes_drt_b200 import ridge
    from oracle import ridge as oridge

    def ridge_fit(WA_re, WA_im, WZ_re, WZ_im, Pen, Lmat, penalty='discrete', nonneg=True, max_iter=20, xtol=1e-3,
                  hl_beta=2.5, lambda_0=1e-2, reg_ord=(0.0, 0.0, 1.0), L1_penalty=0.0, epsilon=1.0, fit_inductance=True,
                  hl_fbeta=None, stop_rule=1, device=None):
        B = WZ_re.shape[0]
        coef, lam, iters, conv = [], [], [], []
        for b in range(B):
            war = (WA_re[b] if WA_re.dim() == 3 else WA_re).numpy()
            wai = (WA_im[b] if WA_im.dim() == 3 else WA_im).numpy()
            c, l, hist, cv = oridge.hyper_loop(war, wai, WZ_re[b].numpy(), WZ_im[b].numpy(), Pen.numpy(),
                                               None if Lmat is None else Lmat.numpy(), penalty, np.asarray(reg_ord),
                                               nonneg, hl_beta, hl_fbeta, lambda_0, L1_penalty, epsilon, xtol, max_iter,
                                               fit_inductance, stop_rule='unchanged' if stop_rule == 1 else 'nan')
            coef.append(c), lam.append(l.copy()), iters.append(len(hist)), conv.append(int(cv))
        return dict(coef=torch.tensor(np.stack(coef)), lam=torch.tensor(np.stack(lam)),
                    iters=torch.tensor(iters, dtype=torch.int32), converged=torch.tensor(conv, dtype=torch.int32))

    def qp_bound(P, q, lb, device=None):
        xs = [oridge.qp_bound(P[b].numpy(), q[b].numpy(), lb.numpy())[0] for b in range(P.shape[0])]
        return torch.tensor(np.stack(xs)), None, None

    def build_M(freq, eps, order, toeplitz, device=None):
        return torch.tensor(om.construct_M(np.asarray(freq, dtype=np.float64), order=order, epsilon=float(eps)))
    monkeypatch.setattr(ridge.capi, 'ridge_fit', ridge_fit)
    monkeypatch.setattr(ridge.capi, 'qp_bound', qp_bound)
    monkeypatch.setattr(ridge.capi, 'build_M', build_M)
    return host


RIDGE_REFERENCE_CASES = {
    'default': dict(), 'huang': dict(preset='Huang'),
    'init_from_ridge': dict(penalty='integral', lambda_0=1, hl_beta=5, weights='modulus'),
    'free_sign': dict(nonneg=False), 'mixed_orders': dict(reg_ord=[0.2, 0.3, 0.5], L1_penalty=0.01),
    'real_part': dict(part='real'), 'imag_part': dict(part='imag', weights='modulus'), 'fbeta': dict(hl_fbeta=0.1),
    'cholesky': dict(penalty='cholesky'), 'cv': dict(lambda_0='cv', cv_lambdas=np.logspace(-6, 0, 7)),
    'ciucci': dict(preset='Ciucci', cv_lambdas=np.logspace(-6, 0, 7)),
}


@pytest.mark.parametrize('case', sorted(RIDGE_REFERENCE_CASES))
def test_shipped_ridge_fit_reproduces_the_reference(case, host_ridge):
    """bayes_drt_b200's ridge_fit (presets, weights, parts and their least-squares offsets, Re-Im cross-validation with
    per-spectrum optima, rescaling) on a batch of two spectra, against the results of the reference's own ridge_fit
    (tests/golden/ridge_reference.npz); the hyper-lambda loop / QP behind the C ABI is served by the oracle here and by
    the CUDA kernels in tests/test_gpu_ridge.py."""
    import warnings
    R = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ridge_reference.npz'))
    S = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'spectra.npz'))
    names = ('ZARC_uniform_0.25', '2ZARC_uniform_0.25')
    Zb = np.stack([S[n + '/Z'] for n in names])
    inv = host_ridge.Inverter()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        # stop_rule='nan': the reference's own loop with an exact QP solver, which the golden vectors were made with
        inv.ridge_fit(S[names[0] + '/freq'], Zb, stop_rule='nan', **RIDGE_REFERENCE_CASES[case])
    tol = 1e-6 if case in ('fbeta', 'ciucci') else 1e-9
    for b, n in enumerate(names):
        key = f'{n}/{case}'
        ref = R[key + '/coef']
        assert np.max(np.abs(inv.distribution_fits['DRT']['coef'][b].numpy() - ref)) <= tol * np.abs(ref).max(), key
        assert float(inv.R_inf[b]) == pytest.approx(float(R[key + '/R_inf']), rel=tol)
        assert abs(float(inv.inductance[b]) - float(R[key + '/inductance'])) <= tol * max(abs(float(R[key + '/inductance'])), 1e-9)
        if key + '/cv_result' in R.files:
            tab = R[key + '/cv_result']
            assert inv._cv_lambda_0[b] == tab[np.argmin(tab[:, 3]), 0]
            for j, k in ((1, 'recv'), (2, 'imcv'), (3, 'totcv')):
                assert np.allclose(inv.cv_result[k][b].numpy(), tab[:, j], rtol=1e-6), (key, k)


def test_per_spectrum_grids_prepare_each_row_like_a_single_fit(host):
    """frequencies [B, Nf]: row b of what fit() hands the solver (grid, scaled spectrum, kernel matrix on that row's own
    default basis) is what a single-spectrum fit of row b hands it -- which is pinned to the reference above."""
    rng = np.random.RandomState(4)
    deltas = np.array([0.0, 0.13, 0.5])
    freq = 10.0 ** (6 - deltas[:, None] - np.arange(81)[None, :] / 10)
    Zb = 1.0 + 1.0 / (1 + (2j * np.pi * freq * 10.0 ** rng.uniform(-4, 0, (3, 1))) ** 0.8) \
        + 0.002 * (rng.randn(3, 81) + 1j * rng.randn(3, 81))
    perm = rng.permutation(81)  # unsorted input: every row is sorted on its own
    inv = host.Inverter()
    with pytest.raises(Abort):
        inv.fit(freq[:, perm], Zb[:, perm], mode='optimize')
    batch = Recorder.last
    assert tuple(batch.A.shape) == (3, 162, 101) and tuple(batch.freq.shape) == (3, 81) and batch.B == 3
    assert tuple(np.shape(inv.distributions['DRT']['tau'])) == (3, 101)
    for b in range(3):
        one = host.Inverter()
        with pytest.raises(Abort):
            one.fit(freq[b], Zb[b], mode='optimize')
        single = Recorder.last
        assert np.array_equal(batch.freq[b].numpy(), single.freq.numpy())
        assert np.allclose(batch.Z[b].numpy(), single.Z[0].numpy(), rtol=1e-14, atol=0)
        assert np.allclose(inv.distributions['DRT']['tau'][b], one.distributions['DRT']['tau'], rtol=1e-15)
        assert np.max(np.abs(batch.A[b].numpy() - single.A.numpy())) <= 1e-13 * np.abs(single.A.numpy()).max()
        assert np.max(np.abs(batch.L.numpy() - single.L.numpy())) <= 1e-12 * np.abs(single.L.numpy()).max()
    with pytest.raises(ValueError):
        host.Inverter().fit(freq, Zb[0])


def test_per_spectrum_grids_reach_the_ridge_solver_row_by_row(host, monkeypatch):
    """ridge_fit (and with it init_from_ridge / outliers='auto' / check_outliers) on frequencies [B, Nf]: row b of the
    weighted augmented system handed to bdrt_ridge_fit is the system of a single-grid fit of row b (pinned to the
    reference above); the penalty matrices are shared."""
    from bayes_drt_b200 import ridge
    rng = np.random.RandomState(5)
    deltas = np.array([0.0, 0.31, 0.77])
    freq = 10.0 ** (6 - deltas[:, None] - np.arange(81)[None, :] / 10)
    Zb = 1.0 + 1.0 / (1 + (2j * np.pi * freq * 10.0 ** rng.uniform(-4, 0, (3, 1))) ** 0.8) \
        + 0.002 * (rng.randn(3, 81) + 1j * rng.randn(3, 81))
    rec = {}

    def ridge_fit(WA_re, WA_im, WZ_re, WZ_im, Pen, Lmat, **o):
        rec.update(WA_re=WA_re, WA_im=WA_im, WZ_re=WZ_re, WZ_im=WZ_im, Pen=Pen, Lmat=Lmat, o=o)
        raise Abort

    def build_M(freq, eps, order, toeplitz, device=None):
        return torch.tensor(om.construct_M(np.asarray(freq, dtype=np.float64), order=order, epsilon=float(eps)))
    monkeypatch.setattr(ridge.capi, 'ridge_fit', ridge_fit)
    monkeypatch.setattr(ridge.capi, 'build_M', build_M)
    for kw in (dict(), dict(preset='Huang')):
        inv = host.Inverter()
        with pytest.raises(Abort):
            inv.ridge_fit(freq, Zb, **kw)
        batch = dict(rec)
        assert tuple(batch['WA_re'].shape) == (3, 81, 103)
        for b in range(3):
            one = host.Inverter()
            with pytest.raises(Abort):
                one.ridge_fit(freq[b], Zb[b], **kw)
            for k in ('WA_re', 'WA_im'):
                single = rec[k] if rec[k].dim() == 2 else rec[k][0]
                assert np.max(np.abs(batch[k][b].numpy() - single.numpy())) <= 1e-13 * np.abs(single.numpy()).max(), k
            for k in ('WZ_re', 'WZ_im'):
                assert np.allclose(batch[k][b].numpy(), rec[k][0].numpy(), rtol=1e-14, atol=0), k
            assert np.max(np.abs(batch['Pen'].numpy() - rec['Pen'].numpy())) <= 1e-12 * np.abs(rec['Pen'].numpy()).max()
