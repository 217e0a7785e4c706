"""Oracle restatement of the reference's per-fit preprocessing and of the Stan programs' log-density (numpy, FP64).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  * Inverter._prep_matrices      inversion.py:2127-2336  (sort, scale, default tau grid, epsilon, A / L)
  * Inverter._scale_Z            inversion.py:2411-2443
  * Inverter._prep_stan_data     inversion.py:1684-1880  (model constants per mode)
  * Stan programs                bayes_drt/stan_model_files/Series_modelcode.txt:1-73,
                                 Series_pos_modelcode.txt:27, Series_outliers_modelcode.txt:1-73,
                                 Series_pos_outliers_modelcode.txt:25
  * Stan transforms / densities  [Stan-upstream 2.19.1]: real<lower=0> theta = exp(u), Jacobian +u (sampling only);
                                 '~' statements drop constant terms.

Pinning.  Data preparation: against the Stan input the reference's own Inverter.fit builds (tests/golden/stan_data.npz,
tests/test_oracle_stan_data.py).  Log-density and gradient: against a mechanical evaluation of the reference's Stan
source text (scripts/stan_subset_interpreter.py reads the *_modelcode.txt files in place; tests/golden/
stan_logdensity.npz, tests/test_oracle_stan_source.py), against an independent literal transcription differentiated by
torch autograd (tests/test_oracle_model.py) and, loosely (~1 % of peak), against the paper's MAP outputs
code_EchemActa/map_results/*.csv.  The Stan / pystan natives themselves are neither in the reference tree nor in this
image, so Stan's own floating-point evaluation is not reproduced bit for bit.
"""
import numpy as np

from . import matrices as om


# ----------------------------------------------------------------------------------------------------------------------
# preprocessing
# ----------------------------------------------------------------------------------------------------------------------
def default_tau(freq):
    """inversion.py:2191-2199: one decade beyond the measured range each side, 10 points per decade."""
    tmin = np.log10(1 / (2 * np.pi * np.max(freq))) - 1
    tmax = np.log10(1 / (2 * np.pi * np.min(freq))) + 1
    num_decades = tmax - tmin
    return np.logspace(tmin, tmax, int(10 * num_decades + 1))


def default_epsilon(tau):
    """inversion.py:2202-2205."""
    return 1.0 / np.mean(np.diff(np.log(tau)))


def z_scale(Z):
    """inversion.py:2437-2441 (series / mixed branch): std(|Z|)/sqrt(Nf/81)."""
    Zmod = np.abs(Z)
    return np.std(Zmod) / np.sqrt(len(Z) / 81)


MODE_CONSTANTS = {
    # inversion.py:1725-1737
    'sample': dict(ups_alpha=1.0, ups_beta=0.1, l0=1.0, l1=1.0, l2=0.75, sigma_out_alpha=5.0),
    'optimize': dict(ups_alpha=0.05, ups_beta=0.1, l0=1.5 * 0.24, l1=1.5 * 0.16, l2=1.5 * 0.08, sigma_out_alpha=2.0),
}


def z_scale_parallel(Z, bc):
    """inversion.py:2417-2434: pure parallel planar DDT fits scale Z so that the scaled admittance has a fixed std
    (14 transmissive / 2.4 blocking)."""
    Ystar_std = {'transmissive': 14.0, 'blocking': 2.4}[bc]
    return Ystar_std * np.sqrt(len(Z) / 81) / np.std(np.abs(1 / Z))


def prep_parallel(freq, Z, info, mode='optimize', sigma_min=0.002, inductance_scale=1.0, scale_Z=True):
    """Stan data of the single-DDT 'Parallel' model (Parallel_modelcode.txt; same constants as Series,
    inversion.py:1714-1754; ``vector<lower=0>[K] x``).  ``info``: the reference's distribution dict
    ({'kernel': 'DDT', 'dist_type': 'parallel', 'symmetry': 'planar', 'bc': ..., 'basis_freq': ...})."""
    freq = np.asarray(freq, dtype=np.float64)
    Z = np.asarray(Z, dtype=np.complex128)
    idx = np.argsort(freq)[::-1]
    freq, Z = freq[idx], Z[idx]
    if not scale_Z:
        zs = 1.0
    elif info['kernel'] == 'DDT' and info.get('symmetry', 'planar') == 'planar':
        zs = z_scale_parallel(Z, info['bc'])
    else:
        zs = z_scale(Z)
    bf = info.get('basis_freq')
    tau = default_tau(freq) if bf is None else 1.0 / (2 * np.pi * np.asarray(bf, dtype=np.float64))
    eps = info.get('epsilon') or default_epsilon(tau)
    kw = dict(tau=tau, epsilon=eps, kernel=info['kernel'], dist_type='parallel',
              symmetry=info.get('symmetry', 'planar'), bc=info.get('bc'), ct=info.get('ct', False),
              k_ct=info.get('k_ct'))
    A_re, A_im = om.construct_A(freq, 'real', **kw), om.construct_A(freq, 'imag', **kw)
    d = prep_series(freq, Z / zs, basis_freq=1 / (2 * np.pi * tau), epsilon=eps, mode=mode, nonneg=True,
                    sigma_min=sigma_min, inductance_scale=inductance_scale, scale_Z=False, A_re=A_re, A_im=A_im)
    d['Z_scale'] = zs
    d['parallel'] = True
    return d


def prep_series(freq, Z, basis_freq=None, epsilon=None, mode='optimize', nonneg=False, outliers=False,
                sigma_min=0.002, inductance_scale=1.0, outlier_lambda=None, scale_Z=True, A_re=None, A_im=None):
    """Build the Stan data of the single-DRT ('Series*') models for one spectrum.

    Returns a dict with the stacked A (2Nf x K), stacked scaled Z, scaled L0/L1/L2 and the scalar constants
    (inversion.py:1714-1754, :1873-1880), plus 'Z_scale', 'tau', 'epsilon', 'freq' (sorted descending, :2138-2141)."""
    freq = np.asarray(freq, dtype=np.float64)
    Z = np.asarray(Z, dtype=np.complex128)
    idx = np.argsort(freq)[::-1]
    freq, Z = freq[idx], Z[idx]
    zs = z_scale(Z) if scale_Z else 1.0
    Zs = Z / zs
    tau = default_tau(freq) if basis_freq is None else 1.0 / (2 * np.pi * np.asarray(basis_freq, dtype=np.float64))
    eps = default_epsilon(tau) if epsilon is None else float(epsilon)
    if A_re is None:
        A_re = om.construct_A(freq, 'real', tau=tau, epsilon=eps)
        A_im = om.construct_A(freq, 'imag', tau=tau, epsilon=eps)
    bf = 1.0 / (2 * np.pi * tau)
    c = MODE_CONSTANTS[mode]
    d = dict(
        Nf=len(freq), K=len(tau), freq=freq, tau=tau, epsilon=eps, Z_scale=zs,
        A=np.concatenate((A_re, A_im)), Z=np.concatenate((Zs.real, Zs.imag)),
        L0=c['l0'] * om.construct_L(bf, tau=tau, epsilon=eps, order=0),
        L1=c['l1'] * om.construct_L(bf, tau=tau, epsilon=eps, order=1),
        L2=c['l2'] * om.construct_L(bf, tau=tau, epsilon=eps, order=2),
        sigma_min=float(sigma_min), ups_alpha=c['ups_alpha'], ups_beta=c['ups_beta'],
        induc_scale=float(inductance_scale), pos=bool(nonneg), outliers=bool(outliers),
        sigma_out_lambda=10.0 if outlier_lambda is None else float(outlier_lambda),  # inversion.py:1708-1712
        sigma_out_alpha=c['sigma_out_alpha'], sigma_out_beta=1.0,
    )
    return d


def n_params(d):
    """D = 2K + 9 (+ 2Nf with the outlier model)  (Series_modelcode.txt:24-36, Series_outliers_modelcode.txt:22-36)."""
    return 2 * d['K'] + 9 + (2 * d['Nf'] if d['outliers'] else 0)


def param_slices(d):
    """Slices of the unconstrained vector in Stan declaration order."""
    K, Nf = d['K'], d['Nf']
    s = {'Rinf_raw': slice(0, 1), 'induc_raw': slice(1, 2), 'x': slice(2, 2 + K),
         'sigma_res_raw': slice(2 + K, 3 + K), 'alpha_prop_raw': slice(3 + K, 4 + K),
         'alpha_re_raw': slice(4 + K, 5 + K), 'alpha_im_raw': slice(5 + K, 6 + K)}
    o = 6 + K
    if d['outliers']:
        s['sigma_out_raw'] = slice(o, o + Nf)
        s['sigma_out_scale'] = slice(o + Nf, o + 2 * Nf)
        o += 2 * Nf
    s['ups_raw'] = slice(o, o + K)
    s['d_strength'] = slice(o + K, o + K + 3)
    return s


def constrain(u, d):
    """Unconstrained vector -> dict of constrained + transformed parameters (the names the reference reads back,
    inversion.py:1229-1276)."""
    sl = param_slices(d)
    K, Nf = d['K'], d['Nf']
    out = {}
    out['Rinf_raw'] = np.exp(u[0])
    out['induc_raw'] = np.exp(u[1])
    out['x'] = np.exp(u[sl['x']]) if d['pos'] else u[sl['x']].copy()
    for nm in ('sigma_res_raw', 'alpha_prop_raw', 'alpha_re_raw', 'alpha_im_raw'):
        out[nm] = np.exp(u[sl[nm]][0])
    out['ups_raw'] = np.exp(u[sl['ups_raw']])
    out['d_strength'] = np.exp(u[sl['d_strength']])
    out['Rinf'] = 100.0 * out['Rinf_raw']
    out['induc'] = out['induc_raw'] * d['induc_scale']
    for nm in ('sigma_res', 'alpha_prop', 'alpha_re', 'alpha_im'):
        out[nm] = 0.05 * out[nm + '_raw']
    zhat = d['A'] @ out['x']
    if d.get('parallel'):  # Parallel_modelcode.txt:46-50: Z_hat_p = 1 / (Y' + i Y'')
        M = zhat[:Nf] ** 2 + zhat[Nf:] ** 2
        zhat = np.concatenate((zhat[:Nf] / M, -zhat[Nf:] / M))
    zhat[:Nf] += out['Rinf']
    zhat[Nf:] += out['induc'] * 2 * np.pi * d['freq']
    out['Z_hat'] = zhat
    s2 = d['sigma_min'] ** 2 + out['sigma_res'] ** 2 + (out['alpha_prop'] * zhat) ** 2 \
        + np.tile((out['alpha_re'] * zhat[:Nf]) ** 2 + (out['alpha_im'] * zhat[Nf:]) ** 2, 2)
    if d['outliers']:
        out['sigma_out_raw'] = np.exp(u[sl['sigma_out_raw']])
        out['sigma_out_scale'] = np.exp(u[sl['sigma_out_scale']])
        out['sigma_out'] = out['sigma_out_raw'] * out['sigma_out_scale'] * 0.05
        s2 = s2 + np.tile(out['sigma_out'] ** 2, 2)
    out['sigma_tot'] = np.sqrt(s2)
    out['ups'] = 0.15 * out['ups_raw']
    return out


# ----------------------------------------------------------------------------------------------------------------------
# log density + analytic gradient
# ----------------------------------------------------------------------------------------------------------------------
def logpost(u, d, jacobian=False, want_grad=True):
    """log p(u | data) up to Stan's dropped constants, and its gradient w.r.t. the unconstrained vector.

    Series_modelcode.txt:37-69 (+ outlier terms Series_outliers_modelcode.txt:45-51, :71-72).  ``jacobian=False`` is
    what StanModel.optimizing maximises, ``jacobian=True`` what StanModel.sampling targets [Stan-upstream]."""
    K, Nf = d['K'], d['Nf']
    sl = param_slices(d)
    A, Z = d['A'], d['Z']
    w = 2 * np.pi * d['freq']
    th = np.exp(u)  # used for every lower=0 parameter
    Rinf_raw, induc_raw = th[0], th[1]
    ux = u[sl['x']]
    x = th[sl['x']] if d['pos'] else ux
    sr_raw, ap_raw, are_raw, aim_raw = th[2 + K:6 + K]
    ups_raw = th[sl['ups_raw']]
    dstr = th[sl['d_strength']]
    sr, ap, are, aim = 0.05 * sr_raw, 0.05 * ap_raw, 0.05 * are_raw, 0.05 * aim_raw

    zhat = A @ x
    par = bool(d.get('parallel'))
    if par:  # Parallel_modelcode.txt:46-50: the distribution contributes an admittance, Z_hat_p = 1 / (Y' + i Y'')
        Yr, Yi = zhat[:Nf].copy(), zhat[Nf:].copy()
        M = Yr ** 2 + Yi ** 2
        zhat = np.concatenate((Yr / M, -Yi / M))
    zhat[:Nf] += 100.0 * Rinf_raw
    zhat[Nf:] += induc_raw * d['induc_scale'] * w
    zre, zim = zhat[:Nf], zhat[Nf:]
    common = (are * zre) ** 2 + (aim * zim) ** 2
    if d['outliers']:
        so_raw = th[sl['sigma_out_raw']]
        so_scale = th[sl['sigma_out_scale']]
        so = 0.05 * so_raw * so_scale
        common = common + so ** 2
    s = d['sigma_min'] ** 2 + sr ** 2 + (ap * zhat) ** 2 + np.tile(common, 2)  # sigma_tot^2
    r = Z - zhat

    a0, a1, a2 = d['L0'] @ x, d['L1'] @ x, d['L2'] @ x
    q2 = dstr[0] * a0 ** 2 + dstr[1] * a1 ** 2 + dstr[2] * a2 ** 2
    ups = 0.15 * ups_raw
    e = 0.5 - 0.25 * (ups[:-2] + ups[2:]) / ups[1:-1]  # dups

    lp = np.sum(-6.0 * np.log(dstr) - 5.0 / dstr)
    lp += np.sum(-(d['ups_alpha'] + 1) * np.log(ups_raw) - d['ups_beta'] / ups_raw)
    lp += -0.5 * (Rinf_raw ** 2 + induc_raw ** 2 + sr_raw ** 2 + ap_raw ** 2 + are_raw ** 2 + aim_raw ** 2)
    lp += np.sum(-0.5 * q2 / ups ** 2 - np.log(ups))
    lp += -0.5 * np.sum(e ** 2)
    lp += np.sum(-0.5 * r ** 2 / s - 0.5 * np.log(s))
    if d['outliers']:
        lp += -d['sigma_out_lambda'] * np.sum(so_raw)
        lp += np.sum(-(d['sigma_out_alpha'] + 1) * np.log(so_scale) - d['sigma_out_beta'] / so_scale)
    lower0 = np.ones(len(u), dtype=bool)
    if not d['pos']:
        lower0[sl['x']] = False
    if jacobian:
        lp += np.sum(u[lower0])
    if not want_grad:
        return lp

    g = 0.5 * r ** 2 / s ** 2 - 0.5 / s  # dlp/ds_i
    G = g[:Nf] + g[Nf:]
    v = r / s + 2 * ap ** 2 * zhat * g
    v[:Nf] += 2 * are ** 2 * zre * G
    v[Nf:] += 2 * aim ** 2 * zim * G
    iu2 = 1.0 / ups ** 2
    vx = v
    if par:  # d lp / d Y from v = d lp / d Z_hat
        c1, c2 = (Yi ** 2 - Yr ** 2) / M ** 2, 2 * Yr * Yi / M ** 2
        vx = np.concatenate((v[:Nf] * c1 + v[Nf:] * c2, -v[:Nf] * c2 + v[Nf:] * c1))
    gx = A.T @ vx - (d['L0'].T @ (dstr[0] * a0 * iu2) + d['L1'].T @ (dstr[1] * a1 * iu2) + d['L2'].T @ (dstr[2] * a2 * iu2))

    grad = np.empty_like(u)
    grad[0] = 100.0 * np.sum(v[:Nf]) - Rinf_raw
    grad[1] = d['induc_scale'] * np.sum(w * v[Nf:]) - induc_raw
    grad[sl['x']] = gx
    grad[2 + K] = 0.05 * 2 * sr * np.sum(g) - sr_raw
    grad[3 + K] = 0.05 * 2 * ap * np.sum(g * zhat ** 2) - ap_raw
    grad[4 + K] = 0.05 * 2 * are * np.sum(G * zre ** 2) - are_raw
    grad[5 + K] = 0.05 * 2 * aim * np.sum(G * zim ** 2) - aim_raw
    if d['outliers']:
        dso = 2 * so * G  # dlp/dsigma_out_n
        grad[sl['sigma_out_raw']] = dso * 0.05 * so_scale - d['sigma_out_lambda']
        grad[sl['sigma_out_scale']] = dso * 0.05 * so_raw - (d['sigma_out_alpha'] + 1) / so_scale \
            + d['sigma_out_beta'] / so_scale ** 2
    gd = np.array([-0.5 * np.sum(a0 ** 2 * iu2), -0.5 * np.sum(a1 ** 2 * iu2), -0.5 * np.sum(a2 ** 2 * iu2)])
    grad[sl['d_strength']] = gd - 6.0 / dstr + 5.0 / dstr ** 2
    gu = q2 / ups ** 3 - 1.0 / ups
    um = ups[1:-1]
    gu[:-2] += e * 0.25 / um
    gu[2:] += e * 0.25 / um
    gu[1:-1] -= e * 0.25 * (ups[:-2] + ups[2:]) / um ** 2
    grad[sl['ups_raw']] = 0.15 * gu - (d['ups_alpha'] + 1) / ups_raw + d['ups_beta'] / ups_raw ** 2
    # chain rule theta = exp(u) for every lower=0 parameter (+1 with the Jacobian)
    grad[lower0] = grad[lower0] * th[lower0] + (1.0 if jacobian else 0.0)
    return lp, grad
