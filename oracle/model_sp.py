"""Oracle restatement of the reference's two-distribution Stan program ``Series-Parallel[_pos]`` (numpy, FP64).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  * Stan program                bayes_drt/stan_model_files/Series-Parallel_modelcode.txt:1-111 and
                                Series-Parallel_pos_modelcode.txt:35 (``vector<lower=0>[Ks] xs``)
  * Inverter._prep_stan_data    inversion.py:1886-1959 (constants per mode; note the parallel L0 multiplier
                                1.5*0.36 in 'optimize', x_sum_invscale 1 ('sample') / 0 ('optimize'), xp_scale from
                                distributions[...]['x_scale'])
  * Inverter._prep_matrices     inversion.py:2127-2336 per distribution; _scale_Z :2437-2441 (mixed model branch)
  * result rescaling            inversion.py:2445-2450 (series coef * Z_scale, parallel coef / Z_scale)

Unconstrained vector, Stan declaration order (Series-Parallel_modelcode.txt:32-49):
  Rinf_raw, induc_raw, xs[Ks], xp_raw[Kp], sigma_res_raw, alpha_prop_raw, alpha_re_raw, alpha_im_raw,
  ups_s_raw[Ks], ups_p_raw[Kp], d0s, d1s, d2s, d0p, d1p, d2p          ->  D = 2 (Ks + Kp) + 12
Every parameter is ``lower=0`` (theta = exp(u)) except ``xs`` in the non-_pos variant.  The transformed parameter
``real<lower=0> x_sum_raw = sum(xs) + sum(xp_raw)`` (:56) is validity-checked by Stan: a negative value rejects the
point (log density -inf); that can only happen in the non-_pos variant.

Pinned to the Stan source text (tests/test_oracle_stan_source.py) and, for Series-Parallel_pos and Series-2Parallel_pos, to
Stan's own output: at the parameter values of the paper's saved MAP fits ``constrain`` reproduces Stan's transformed
parameters to 1e-15 (tests/test_oracle_stan_map.py; also fixes the by-name ordering of two parallel distributions).
"""
import numpy as np

from . import matrices as om
from .model import default_epsilon, default_tau, z_scale

# inversion.py:1913-1930
MODE_CONSTANTS_SP = {
    'sample': dict(ups_alpha=1.0, ups_beta=0.1, ls=(1.0, 1.0, 0.75), lp=(1.0, 1.0, 0.75), x_sum_invscale=1.0),
    'optimize': dict(ups_alpha=0.05, ups_beta=0.1, ls=(1.5 * 0.24, 1.5 * 0.16, 1.5 * 0.08),
                     lp=(1.5 * 0.36, 1.5 * 0.16, 1.5 * 0.08), x_sum_invscale=0.0),
}


def _dist_matrices(freq, info):
    """A_re, A_im, L0, L1, L2, tau, epsilon of one distribution (inversion.py:2191-2209, :2250-2271, :2301-2307)."""
    bf = info.get('basis_freq')
    tau = default_tau(freq) if bf is None else 1.0 / (2 * np.pi * np.asarray(bf, dtype=np.float64))
    eps = info.get('epsilon') or default_epsilon(tau)
    kw = dict(tau=tau, epsilon=eps, kernel=info['kernel'], dist_type=info['dist_type'],
              symmetry=info.get('symmetry', 'planar'), bc=info.get('bc'), ct=info.get('ct', False),
              k_ct=info.get('k_ct'))
    A_re = om.construct_A(freq, 'real', **kw)
    A_im = om.construct_A(freq, 'imag', **kw)
    bfr = 1.0 / (2 * np.pi * tau)
    L = [om.construct_L(bfr, tau=tau, epsilon=eps, order=o) for o in (0, 1, 2)]
    return A_re, A_im, L, tau, eps


def prep_series_2parallel(freq, Z, ser, par1, par2, mode='optimize', nonneg=True, sigma_min=0.002,
                          inductance_scale=1.0, scale_Z=True):
    """Stan data of the Series-2Parallel model (Series-2Parallel[_pos]_modelcode.txt; constants inversion.py:1961-2049:
    as Series-Parallel for both parallel distributions, x_sum_invscale 0.1 ('sample') / 0 ('optimize')).
    The reference orders the parallel distributions by *sorted name* (:1963); the caller passes them in that order.
    Unconstrained vector: Rinf_raw, induc_raw, xs, xp1_raw, xp2_raw, 4 error parameters, ups_s_raw, ups_p1_raw,
    ups_p2_raw, d_s(3), d_p1(3), d_p2(3)   ->  D = 2 (Ks + Kp1 + Kp2) + 15."""
    d = prep_series_parallel(freq, Z, ser, par1, mode=mode, nonneg=nonneg, sigma_min=sigma_min,
                             inductance_scale=inductance_scale, scale_Z=scale_Z)
    par2 = dict(par2, dist_type='parallel')
    A_re, A_im, L, tau, eps = _dist_matrices(d['freq'], par2)
    c = MODE_CONSTANTS_SP[mode]
    d.update(Kp2=len(tau), tau_p2=tau, eps_p2=eps, Ap2=np.concatenate((A_re, A_im)),
             Lp2=[c['lp'][o] * L[o] for o in range(3)], xp2_scale=float(par2.get('x_scale', 1)),
             x_sum_invscale=0.1 if mode == 'sample' else 0.0)
    return d


def _pars(d):
    """[(K, A, L, x_scale)] of the parallel distributions of a data dict (one for Series-Parallel, two for Series-2Parallel)"""
    out = [(d['Kp'], d['Ap'], d['Lp'], d['xp_scale'])]
    if 'Kp2' in d:
        out.append((d['Kp2'], d['Ap2'], d['Lp2'], d['xp2_scale']))
    return out


def prep_series_parallel(freq, Z, ser, par, mode='optimize', nonneg=True, sigma_min=0.002, inductance_scale=1.0,
                         scale_Z=True):
    """Stan data of the Series-Parallel model for one spectrum.  ``ser`` / ``par``: the reference's distribution
    info dicts ({'kernel': 'DRT', ...} / {'kernel': 'DDT', 'dist_type': 'parallel', 'symmetry': 'planar',
    'bc': 'transmissive', 'x_scale': 0.8, ...})."""
    freq = np.asarray(freq, dtype=np.float64)
    Z = np.asarray(Z, dtype=np.complex128)
    idx = np.argsort(freq)[::-1]
    freq, Z = freq[idx], Z[idx]
    zs = z_scale(Z) if scale_Z else 1.0
    Zs = Z / zs
    ser = dict(ser, dist_type='series')
    par = dict(par, dist_type='parallel')
    As_re, As_im, Ls, tau_s, eps_s = _dist_matrices(freq, ser)
    Ap_re, Ap_im, Lp, tau_p, eps_p = _dist_matrices(freq, par)
    c = MODE_CONSTANTS_SP[mode]
    return dict(
        Nf=len(freq), Ks=len(tau_s), Kp=len(tau_p), freq=freq, tau_s=tau_s, tau_p=tau_p, eps_s=eps_s, eps_p=eps_p,
        Z_scale=zs, As=np.concatenate((As_re, As_im)), Ap=np.concatenate((Ap_re, Ap_im)),
        Z=np.concatenate((Zs.real, Zs.imag)),
        Ls=[c['ls'][o] * Ls[o] for o in range(3)], Lp=[c['lp'][o] * Lp[o] for o in range(3)],
        sigma_min=float(sigma_min), ups_alpha=c['ups_alpha'], ups_beta=c['ups_beta'],
        induc_scale=float(inductance_scale), x_sum_invscale=c['x_sum_invscale'],
        xp_scale=float(par.get('x_scale', 1)), pos=bool(nonneg))


def n_params(d):
    P = _pars(d)
    return 2 * (d['Ks'] + sum(p[0] for p in P)) + 6 + 3 * (1 + len(P))


def param_slices(d):
    Ks = d['Ks']
    Kp = [p[0] for p in _pars(d)]
    o = 2
    s = {'Rinf_raw': slice(0, 1), 'induc_raw': slice(1, 2), 'xs': slice(o, o + Ks)}
    o += Ks
    s['xp'] = []
    for k in Kp:
        s['xp'].append(slice(o, o + k))
        o += k
    s['xp_raw'] = s['xp'][0]
    s['err'] = slice(o, o + 4)
    o += 4
    s['ups_s_raw'] = slice(o, o + Ks)
    o += Ks
    s['ups_p'] = []
    for k in Kp:
        s['ups_p'].append(slice(o, o + k))
        o += k
    s['ups_p_raw'] = s['ups_p'][0]
    s['ds'] = slice(o, o + 3)
    o += 3
    s['dps'] = []
    for k in Kp:
        s['dps'].append(slice(o, o + 3))
        o += 3
    s['dp'] = s['dps'][0]
    return s


def _prior_block(x, L, dstr, ups):
    """sum_k [-1/2 q_k^2/ups_k^2 - log ups_k] - 1/2 sum dups^2 and its derivatives w.r.t. x, d, ups."""
    a = [L[j] @ x for j in range(3)]
    q2 = sum(dstr[j] * a[j] ** 2 for j in range(3))
    iu2 = 1.0 / ups ** 2
    e = 0.5 - 0.25 * (ups[:-2] + ups[2:]) / ups[1:-1]
    lp = np.sum(-0.5 * q2 * iu2 - np.log(ups)) - 0.5 * np.sum(e ** 2)
    gx = -sum(L[j].T @ (dstr[j] * a[j] * iu2) for j in range(3))
    gd = np.array([-0.5 * np.sum(a[j] ** 2 * iu2) for j in range(3)])
    gu = q2 / ups ** 3 - 1.0 / ups
    um = ups[1:-1]
    gu[:-2] += e * 0.25 / um
    gu[2:] += e * 0.25 / um
    gu[1:-1] -= e * 0.25 * (ups[:-2] + ups[2:]) / um ** 2
    return lp, gx, gd, gu


def logpost(u, d, jacobian=False, want_grad=True):
    """log p(u | data) up to Stan's dropped constants and its gradient (Series-Parallel_modelcode.txt:50-110,
    Series-2Parallel_modelcode.txt)."""
    Ks, Nf = d['Ks'], d['Nf']
    P = _pars(d)
    sl = param_slices(d)
    w = 2 * np.pi * d['freq']
    th = np.exp(u)
    Rinf_raw, induc_raw = th[0], th[1]
    xs = th[sl['xs']] if d['pos'] else u[sl['xs']]
    xp_raw = [th[s_] for s_ in sl['xp']]
    sr_raw, ap_raw, are_raw, aim_raw = th[sl['err']]
    ups_s_raw = th[sl['ups_s_raw']]
    ups_p_raw = [th[s_] for s_ in sl['ups_p']]
    ds = th[sl['ds']]
    dps = [th[s_] for s_ in sl['dps']]
    sr, ap, are, aim = 0.05 * sr_raw, 0.05 * ap_raw, 0.05 * are_raw, 0.05 * aim_raw
    inv = d['x_sum_invscale']

    x_sum_raw = np.sum(xs) + sum(np.sum(x) for x in xp_raw)
    if x_sum_raw < 0:  # real<lower=0> x_sum_raw (:56): Stan rejects the point
        return (-np.inf, np.full_like(u, np.nan)) if want_grad else -np.inf
    x_sum = x_sum_raw * inv
    zhat = d['As'] @ xs
    Ys = []
    for (Kp, Ap, Lp, xsc), xr in zip(P, xp_raw):
        Y = Ap @ (xr * xsc)
        Yr, Yi = Y[:Nf], Y[Nf:]
        M = Yr ** 2 + Yi ** 2
        zhat[:Nf] += Yr / M
        zhat[Nf:] += -Yi / M
        Ys.append((Yr, Yi, M))
    zhat[:Nf] += 100.0 * Rinf_raw
    zhat[Nf:] += induc_raw * d['induc_scale'] * w
    zre, zim = zhat[:Nf], zhat[Nf:]
    common = (are * zre) ** 2 + (aim * zim) ** 2
    s = d['sigma_min'] ** 2 + sr ** 2 + (ap * zhat) ** 2 + np.tile(common, 2)
    r = d['Z'] - zhat
    ups_s = 0.15 * ups_s_raw
    lps, gxs, gds, gus = _prior_block(xs, d['Ls'], ds, ups_s)
    pri = [_prior_block(xr, Lp, dp, 0.15 * ur) for (Kp, Ap, Lp, xsc), xr, dp, ur in zip(P, xp_raw, dps, ups_p_raw)]

    lp = np.sum(-6.0 * np.log(ds) - 5.0 / ds) + sum(np.sum(-6.0 * np.log(dp) - 5.0 / dp) for dp in dps)
    lp += -0.5 * x_sum ** 2
    for ur in [ups_s_raw] + ups_p_raw:
        lp += np.sum(-(d['ups_alpha'] + 1) * np.log(ur) - d['ups_beta'] / ur)
    lp += -0.5 * (Rinf_raw ** 2 + induc_raw ** 2 + sr_raw ** 2 + ap_raw ** 2 + are_raw ** 2 + aim_raw ** 2)
    lp += lps + sum(p_[0] for p_ in pri)
    lp += np.sum(-0.5 * r ** 2 / s - 0.5 * np.log(s))
    lower0 = np.ones(len(u), dtype=bool)
    if not d['pos']:
        lower0[sl['xs']] = False
    if jacobian:
        lp += np.sum(u[lower0])
    if not want_grad:
        return lp

    g = 0.5 * r ** 2 / s ** 2 - 0.5 / s
    G = g[:Nf] + g[Nf:]
    v = r / s + 2 * ap ** 2 * zhat * g
    v[:Nf] += 2 * are ** 2 * zre * G
    v[Nf:] += 2 * aim ** 2 * zim * G
    vr, vi = v[:Nf], v[Nf:]

    grad = np.empty_like(u)
    grad[0] = 100.0 * np.sum(vr) - Rinf_raw
    grad[1] = d['induc_scale'] * np.sum(w * vi) - induc_raw
    grad[sl['xs']] = d['As'].T @ v + gxs - x_sum * inv
    for i, ((Kp, Ap, Lp, xsc), (Yr, Yi, M)) in enumerate(zip(P, Ys)):
        # Z_p = 1 / (Y' + i Y''):  d lp / d Y from v = d lp / d Z_hat
        c1, c2 = (Yi ** 2 - Yr ** 2) / M ** 2, 2 * Yr * Yi / M ** 2
        gY = np.concatenate((vr * c1 + vi * c2, -vr * c2 + vi * c1))
        grad[sl['xp'][i]] = xsc * (Ap.T @ gY) + pri[i][1] - x_sum * inv
        ur = ups_p_raw[i]
        grad[sl['ups_p'][i]] = 0.15 * pri[i][3] - (d['ups_alpha'] + 1) / ur + d['ups_beta'] / ur ** 2
        grad[sl['dps'][i]] = pri[i][2] - 6.0 / dps[i] + 5.0 / dps[i] ** 2
    o = sl['err'].start
    grad[o] = 0.05 * 2 * sr * np.sum(g) - sr_raw
    grad[o + 1] = 0.05 * 2 * ap * np.sum(g * zhat ** 2) - ap_raw
    grad[o + 2] = 0.05 * 2 * are * np.sum(G * zre ** 2) - are_raw
    grad[o + 3] = 0.05 * 2 * aim * np.sum(G * zim ** 2) - aim_raw
    grad[sl['ups_s_raw']] = 0.15 * gus - (d['ups_alpha'] + 1) / ups_s_raw + d['ups_beta'] / ups_s_raw ** 2
    grad[sl['ds']] = gds - 6.0 / ds + 5.0 / ds ** 2
    grad[lower0] = grad[lower0] * th[lower0] + (1.0 if jacobian else 0.0)
    return lp, grad


def constrain(u, d):
    """The names the reference reads back (inversion.py:1236-1258): xs, xp (xp1, xp2), Rinf, induc, sigma_*, alpha_*."""
    Nf = d['Nf']
    P = _pars(d)
    sl = param_slices(d)
    th = np.exp(u)
    out = {'xs': th[sl['xs']] if d['pos'] else u[sl['xs']].copy(), 'Rinf': 100.0 * th[0], 'induc': th[1] * d['induc_scale']}
    for i, nm in enumerate(('sigma_res', 'alpha_prop', 'alpha_re', 'alpha_im')):
        out[nm] = 0.05 * th[sl['err']][i]
    zhat = d['As'] @ out['xs']
    for i, (Kp, Ap, Lp, xsc) in enumerate(P):
        xp = th[sl['xp'][i]] * xsc
        out['xp' if len(P) == 1 else f'xp{i + 1}'] = xp
        Y = Ap @ xp
        M = Y[:Nf] ** 2 + Y[Nf:] ** 2
        zhat[:Nf] += Y[:Nf] / M
        zhat[Nf:] += -Y[Nf:] / M
    zhat[:Nf] += out['Rinf']
    zhat[Nf:] += out['induc'] * 2 * np.pi * d['freq']
    out['Z_hat'] = zhat
    out['sigma_tot'] = np.sqrt(d['sigma_min'] ** 2 + out['sigma_res'] ** 2 + (out['alpha_prop'] * zhat) ** 2
                               + np.tile((out['alpha_re'] * zhat[:Nf]) ** 2 + (out['alpha_im'] * zhat[Nf:]) ** 2, 2))
    return out
