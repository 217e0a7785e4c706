"""CPU: the oracle against REAL Stan output -- the MAP fits the reference's authors saved for their paper
(code_EchemActa/map_results/obj_*.pkl: every parameter and transformed parameter ``StanModel.optimizing`` returned, with
the matrices the fit used; fixture tests/golden/stan_map.npz, made by scripts/make_golden_stan_map.py).  pystan cannot be
installed here, so these are the only numbers in the tree that Stan's own compiled programs computed.

  matrices    the oracle's A', A'', L0, L1, L2 for the stored grids / settings == the matrices the reference used
              (through seeded random projections), 1e-10
  forward     at Stan's own parameter values the oracle reproduces Stan's transformed parameters Z_hat, sigma_tot, q
              (qs, qp), ups, dups -- Series[_pos]_modelcode.txt:38-54, Series-Parallel_pos_modelcode.txt:50-90 -- to 1e-10:
              the model arithmetic, the constants handed to Stan (L scalings per mode, 0.05 / 0.15 / 100 factors, x_scale)
              and the unconstrained <-> constrained maps are Stan's
  optimum     Stan's end points are near-stationary points of the oracle's log density (max|grad| orders of magnitude
              below a random start), and where Stan converged tightly its optimum is the oracle's (Newton from Stan's
              point moves x by < 1e-2 of its peak and lp by < 0.1)

(What this cannot pin is the optimiser's path: the files hold end points only.)"""
import json
import os

import numpy as np
import pytest

from oracle import matrices as om, model as omod, model_sp as osp, newton as onew

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_G = None
PROJ_SEED = 20201113  # scripts/make_golden_stan_map.py


def G():
    global _G
    if _G is None:
        with np.load(os.path.join(ROOT, 'tests', 'golden', 'stan_map.npz')) as g:
            _G = {k: g[k] for k in g.files}
    return _G


def proj_vectors(n_rows, n_cols):
    rng = np.random.RandomState(PROJ_SEED + 1000 * n_rows + n_cols)
    return rng.standard_normal(n_rows), rng.standard_normal(n_cols)


NAMES = [str(n) for n in np.load(os.path.join(ROOT, 'tests', 'golden', 'stan_map.npz'))['names']]
# (added after the last GPU session of the round: checked against the oracle here, not yet run through the CUDA test)
CPU_ONLY = ['LIB_data_DRT-TpDDT', 'LIB_data_qtr_DRT-TpDDT']


def _info(g, p, dd):
    i = dict(kernel=dd['kernel'], dist_type=dd['dist_type'], basis_freq=1 / (2 * np.pi * g[p + 'tau/' + dd['name']]),
             epsilon=dd['epsilon'])
    if dd['kernel'] == 'DDT':
        i.update(symmetry=dd['symmetry'], bc=dd['bc'], ct=dd['ct'], x_scale=dd['x_scale'])
    return i


def build(name):
    """-> (oracle data dict, Stan's point as the oracle's unconstrained vector, Stan's result, module, meta)"""
    g = G()
    p = name + '/'
    meta = json.loads(str(g[p + 'meta']))
    model = str(g[p + 'model'])
    freq, Z = g[p + 'freq'], g[p + 'Z']
    S = {k[len(p) + 5:]: g[k] for k in g if k.startswith(p + 'stan/')}
    dists = meta['dists']
    induc = S['induc'] if 'induc' in S else S['induc_raw']  # that version declared induc itself (induc_scale = 1)
    if model.startswith('Series_'):
        dd = dists[0]
        d = omod.prep_series(freq, Z, basis_freq=1 / (2 * np.pi * g[p + 'tau/' + dd['name']]), epsilon=dd['epsilon'],
                             mode='optimize', nonneg='pos' in model, sigma_min=meta['sigma_min'])
        sl = omod.param_slices(d)
        u = np.zeros(omod.n_params(d))
        u[0], u[1] = np.log(S['Rinf_raw']), np.log(induc)
        u[sl['x']] = np.log(S['x']) if d['pos'] else S['x']
        for nm in ('sigma_res_raw', 'alpha_prop_raw', 'alpha_re_raw', 'alpha_im_raw'):
            u[sl[nm]] = np.log(S[nm])
        u[sl['ups_raw']] = np.log(S['ups_raw'])
        u[sl['d_strength']] = np.log([S['d0_strength'], S['d1_strength'], S['d2_strength']])
        return d, u, S, omod, meta
    ser = [x for x in dists if x['dist_type'] == 'series'][0]
    # (two parallel distributions reach Stan ordered by name, inversion.py:1963: BP-DDT is xp1, TP-DDT xp2)
    pars = sorted((x for x in dists if x['dist_type'] == 'parallel'), key=lambda x: x['name'])
    # x_scale is data handed to Stan (xp = xp_raw * x_scale); two of the saved objects do not carry the value their fit
    # used, so it is read off Stan's own output where both vectors are there
    for i, pr in enumerate(pars):
        q = 'xp' if len(pars) == 1 else f'xp{i + 1}'
        if q in S and q + '_raw' in S:
            ratio = S[q] / S[q + '_raw']
            assert np.ptp(ratio) <= 1e-12 * ratio[0]
            pr['x_scale'] = float(np.median(ratio))
    if len(pars) == 1:
        d = osp.prep_series_parallel(freq, Z, _info(g, p, ser), _info(g, p, pars[0]), mode='optimize', nonneg=True,
                                     sigma_min=meta['sigma_min'])
        xr, ur, dn = ['xp_raw'], ['ups_p_raw'], ['p']
    else:
        d = osp.prep_series_2parallel(freq, Z, _info(g, p, ser), _info(g, p, pars[0]), _info(g, p, pars[1]),
                                      mode='optimize', nonneg=True, sigma_min=meta['sigma_min'])
        xr, ur, dn = ['xp1_raw', 'xp2_raw'], ['ups_p1_raw', 'ups_p2_raw'], ['p1', 'p2']
    meta['pars'] = pars
    sl = osp.param_slices(d)
    u = np.zeros(osp.n_params(d))
    u[0], u[1] = np.log(S['Rinf_raw']), np.log(induc)
    u[sl['xs']] = np.log(S['xs'])
    for i, nm in enumerate(xr):  # (one file is from a program variant without x_scale: xp is the parameter itself)
        u[sl['xp'][i]] = np.log(S[nm] if nm in S else S[nm[:-4]] / pars[i]['x_scale'])
    u[sl['err']] = np.log([S['sigma_res_raw'], S['alpha_prop_raw'], S['alpha_re_raw'], S['alpha_im_raw']])
    u[sl['ups_s_raw']] = np.log(S['ups_s_raw'])
    for i, nm in enumerate(ur):
        u[sl['ups_p'][i]] = np.log(S[nm])
    u[sl['ds']] = np.log([S['d0s_strength'], S['d1s_strength'], S['d2s_strength']])
    for i, q in enumerate(dn):
        u[sl['dps'][i]] = np.log([S[f'd0{q}_strength'], S[f'd1{q}_strength'], S[f'd2{q}_strength']])
    return d, u, S, osp, meta


def _close(a, b, tol):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) <= tol * np.max(np.abs(b))


@pytest.mark.parametrize('name', NAMES)
def test_matrices_are_the_ones_the_reference_used(name):
    g = G()
    p = name + '/'
    meta = json.loads(str(g[p + 'meta']))
    freq = g[p + 'freq']
    for dd in meta['dists']:
        tau = g[p + 'tau/' + dd['name']]
        kw = dict(tau=tau, epsilon=dd['epsilon'], kernel=dd['kernel'], dist_type=dd['dist_type'],
                  symmetry=dd['symmetry'] or 'planar', bc=dd['bc'], ct=dd['ct'])
        mats = {'A_re': om.construct_A(freq, 'real', **kw), 'A_im': om.construct_A(freq, 'imag', **kw)}
        bf = 1 / (2 * np.pi * tau)
        for o in (0, 1, 2):
            mats[f'L{o}'] = om.construct_L(bf, tau=tau, epsilon=dd['epsilon'], order=o)
        for mn, M in mats.items():
            l, r = proj_vectors(*M.shape)
            assert _close(M @ r, g[p + f'proj/{dd["name"]}/{mn}/r'], 1e-10), (name, dd['name'], mn)
            assert _close(l @ M, g[p + f'proj/{dd["name"]}/{mn}/l'], 1e-10), (name, dd['name'], mn)


def _q(x, L, dstr):
    return np.sqrt(dstr[0] * (L[0] @ x) ** 2 + dstr[1] * (L[1] @ x) ** 2 + dstr[2] * (L[2] @ x) ** 2)


def _dups(ups):
    return 0.5 * (ups[1:-1] - 0.5 * (ups[:-2] + ups[2:])) / ups[1:-1]


@pytest.mark.parametrize('name', NAMES)
def test_forward_model_reproduces_stans_transformed_parameters(name):
    d, u, S, mod, meta = build(name)
    assert abs(d['Z_scale'] - meta['Z_scale']) <= 1e-13 * meta['Z_scale']
    c = mod.constrain(u, d)
    assert _close(c['Z_hat'], S['Z_hat'], 1e-10), name
    assert np.max(np.abs(c['sigma_tot'] - S['sigma_tot']) / S['sigma_tot']) <= 1e-10, name
    assert abs(c['Rinf'] - S['Rinf']) <= 1e-12 * S['Rinf']
    if mod is omod:
        dstr = [S['d0_strength'], S['d1_strength'], S['d2_strength']]
        assert _close(_q(c['x'], [d['L0'], d['L1'], d['L2']], dstr), S['q'], 1e-10)
        assert _close(c['ups'], S['ups'], 1e-13)
        assert _close(_dups(c['ups']), S['dups'], 1e-10)
    else:
        assert _close(_q(c['xs'], d['Ls'], [S['d0s_strength'], S['d1s_strength'], S['d2s_strength']]), S['qs'], 1e-10)
        assert _close(_dups(0.15 * S['ups_s_raw']), S['dups_s'], 1e-10)
        P = osp._pars(d)
        for i, (Kp, Ap, Lp, xsc) in enumerate(P):
            q = 'p' if len(P) == 1 else f'p{i + 1}'
            xp = c['xp' if len(P) == 1 else f'xp{i + 1}']
            assert _close(xp, S['x' + q], 1e-12)
            # q_p is built from the raw (unscaled) coefficients (Series-Parallel_pos_modelcode.txt:58-60)
            assert _close(_q(xp / xsc, Lp, [S[f'd0{q}_strength'], S[f'd1{q}_strength'], S[f'd2{q}_strength']]),
                          S['q' + q], 1e-10)
        xs_sum = np.sum(c['xs']) + sum(np.sum(c['xp' if len(P) == 1 else f'xp{i + 1}'] / P[i][3]) for i in range(len(P)))
        assert abs(xs_sum - S['x_sum_raw']) <= 1e-10 * abs(S['x_sum_raw'])


# fits whose end point Stan converged tightly (its last objective decrements are below 1e-2 of a log-density unit)
TIGHT = ['RC-ZARC_Macdonald_2.5', 'RC-ZARC_uniform_2.5', 'RC-ZARC_uniform_1.0', 'trunc_Macdonald_1.0']


@pytest.mark.parametrize('name', [n for n in NAMES if n.startswith(('RC-ZARC', 'trunc'))])
def test_stans_end_point_is_near_stationary_for_the_oracle_density(name):
    d, u, S, mod, meta = build(name)
    lp, g = mod.logpost(u, d)
    assert np.isfinite(lp)
    rng = np.random.RandomState(0)
    g_rand = [np.max(np.abs(mod.logpost(rng.uniform(-2, 2, len(u)), d)[1])) for _ in range(5)]
    assert np.max(np.abs(g)) <= 1e-3 * np.median(g_rand), (np.max(np.abs(g)), g_rand)
    if name in TIGHT:
        def func(z):
            l, gg = mod.logpost(z, d)
            return None if not np.isfinite(l) else (-l, -gg)
        with np.errstate(all='ignore'):
            pz = onew.polish(func, u, max_iter=120)
        assert pz['gnorm'] < 1e-8
        xs, xo = mod.constrain(u, d)['x'], mod.constrain(pz['x'], d)['x']
        assert 0.0 <= -pz['f'] - lp <= 0.1, -pz['f'] - lp
        assert np.max(np.abs(xs - xo)) <= 1e-2 * np.max(np.abs(xo))


@pytest.mark.parametrize('name', ['RC-ZARC_uniform_2.5', 'RC-ZARC_uniform_1.0'])
def test_stan_semantics_lbfgs_from_random_starts_finds_stans_optimum(name):
    """End to end on the data of a saved fit: the restatement of Stan's L-BFGS (oracle/lbfgs.py, the statement
    csrc/lbfgs_kernel.cuh implements) from Stan-style random starts U(-2, 2) ends where pystan ended -- within 0.05 of
    its log density and 5e-3 of the peak in x, i.e. within Stan's own termination accuracy -- for at least one of three
    starts.  (The posterior is multimodal: other starts, like other seeds of pystan, end in other modes; for some of the
    saved fits every random start here finds a mode with a HIGHER density than the saved one.)"""
    from oracle import lbfgs as olb
    d, u, S, mod, meta = build(name)
    lp_stan = mod.logpost(u, d)[0]
    x_stan = mod.constrain(u, d)['x']

    def func(z):
        l, gg = mod.logpost(z, d)
        return None if (not np.isfinite(l) or not np.all(np.isfinite(gg))) else (-l, -gg)
    best = None
    for seed in range(3):
        with np.errstate(all='ignore'):
            r = olb.minimize(func, np.random.RandomState(seed).uniform(-2, 2, len(u)), max_iter=50000)
        dist = np.max(np.abs(mod.constrain(r['x'], d)['x'] - x_stan)) / np.max(np.abs(x_stan))
        if best is None or dist < best[1]:
            best = (abs(-r['f'] - lp_stan), dist)
    assert best[0] <= 0.05 and best[1] <= 5e-3, best
