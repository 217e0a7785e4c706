"""Line-by-line torch transcription of the Stan programs, differentiated by autograd.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Independent second statement of the density used to check the
hand-derived gradient in oracle/model.py: every line below names the Stan line it transcribes
(bayes_drt/stan_model_files/Series_modelcode.txt, Series_outliers_modelcode.txt).
"""
import math

import torch


def _inv_gamma_lpdf(y, a, b):
    # Stan: inv_gamma_lpdf = a log b - lgamma(a) - (a+1) log y - b / y ; '~' drops the first two (data-only) terms
    return torch.sum(-(a + 1) * torch.log(y) - b / y)


def _normal_lpdf(y, mu, sigma):
    # '~' drops -0.5 log(2 pi)
    return torch.sum(-torch.log(sigma) - 0.5 * ((y - mu) / sigma) ** 2)


def _std_normal_lpdf(y):
    return torch.sum(-0.5 * y ** 2)


def logpost_literal(u, d, jacobian=False):
    """u: torch float64 tensor (requires_grad ok); d: dict from oracle.model.prep_series."""
    K, Nf = d['K'], d['Nf']
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    A, Z, freq = t(d['A']), t(d['Z']), t(d['freq'])
    L0, L1, L2 = t(d['L0']), t(d['L1']), t(d['L2'])
    N = 2 * Nf
    # transformed data (Series_modelcode.txt:18-23)
    Rinf_vec = torch.cat((torch.ones(Nf, dtype=torch.float64), torch.zeros(Nf, dtype=torch.float64)))
    induc_vec = torch.cat((torch.zeros(Nf, dtype=torch.float64), 2 * math.pi * freq))
    # parameters (Series_modelcode.txt:24-36): lower=0 -> exp transform, log-Jacobian = u
    pos = 0
    logjac = torch.zeros((), dtype=torch.float64)

    def take(n, lower0):
        nonlocal pos, logjac
        raw = u[pos:pos + n]
        pos += n
        if lower0:
            logjac = logjac + raw.sum()
            return torch.exp(raw)
        return raw

    Rinf_raw = take(1, True)[0]
    induc_raw = take(1, True)[0]
    x = take(K, d['pos'])
    sigma_res_raw = take(1, True)[0]
    alpha_prop_raw = take(1, True)[0]
    alpha_re_raw = take(1, True)[0]
    alpha_im_raw = take(1, True)[0]
    if d['outliers']:
        sigma_out_raw = take(Nf, True)
        sigma_out_scale = take(Nf, True)
    ups_raw = take(K, True)
    d0 = take(1, True)[0]
    d1 = take(1, True)[0]
    d2 = take(1, True)[0]
    # transformed parameters (Series_modelcode.txt:37-54)
    Rinf = Rinf_raw * 100
    induc = induc_raw * d['induc_scale']
    q = torch.sqrt(d0 * (L0 @ x) ** 2 + d1 * (L1 @ x) ** 2 + d2 * (L2 @ x) ** 2)
    sigma_res = sigma_res_raw * 0.05
    alpha_prop = alpha_prop_raw * 0.05
    alpha_re = alpha_re_raw * 0.05
    alpha_im = alpha_im_raw * 0.05
    if d.get('parallel'):  # Parallel_modelcode.txt:46-50
        Y_hat = A @ x
        Y_hat_re, Y_hat_im = Y_hat[:Nf], Y_hat[Nf:]
        Z_hat_p = torch.cat((Y_hat_re / (Y_hat_re ** 2 + Y_hat_im ** 2), -Y_hat_im / (Y_hat_re ** 2 + Y_hat_im ** 2)))
        Z_hat = Z_hat_p + Rinf * Rinf_vec + induc * induc_vec
    else:
        Z_hat = A @ x + Rinf * Rinf_vec + induc * induc_vec
    Z_hat_re = torch.cat((Z_hat[:Nf], Z_hat[:Nf]))
    Z_hat_im = torch.cat((Z_hat[Nf:], Z_hat[Nf:]))
    var = d['sigma_min'] ** 2 + sigma_res ** 2 + (alpha_prop * Z_hat) ** 2 + (alpha_re * Z_hat_re) ** 2 \
        + (alpha_im * Z_hat_im) ** 2
    if d['outliers']:
        sigma_out = sigma_out_raw * sigma_out_scale * 0.05  # Series_outliers_modelcode.txt:45
        var = var + torch.cat((sigma_out, sigma_out)) ** 2  # :49-51
    sigma_tot = torch.sqrt(var)
    ups = ups_raw * 0.15
    dups = 0.5 * (ups[1:-1] - 0.5 * (ups[:-2] + ups[2:])) / ups[1:-1]
    # model (Series_modelcode.txt:55-69)
    lp = _inv_gamma_lpdf(d0, 5.0, 5.0) + _inv_gamma_lpdf(d1, 5.0, 5.0) + _inv_gamma_lpdf(d2, 5.0, 5.0)
    lp = lp + _inv_gamma_lpdf(ups_raw, d['ups_alpha'], d['ups_beta'])
    lp = lp + _std_normal_lpdf(Rinf_raw) + _std_normal_lpdf(induc_raw)
    lp = lp + _normal_lpdf(q, 0.0, ups)
    lp = lp + _std_normal_lpdf(dups)
    lp = lp + _normal_lpdf(Z, Z_hat, sigma_tot)
    lp = lp + _std_normal_lpdf(sigma_res_raw) + _std_normal_lpdf(alpha_prop_raw) + _std_normal_lpdf(alpha_re_raw) \
        + _std_normal_lpdf(alpha_im_raw)
    if d['outliers']:
        lp = lp + torch.sum(-d['sigma_out_lambda'] * sigma_out_raw)  # exponential(lambda), log(lambda) dropped
        lp = lp + _inv_gamma_lpdf(sigma_out_scale, d['sigma_out_alpha'], d['sigma_out_beta'])
    if jacobian:
        lp = lp + logjac
    assert pos == u.numel()
    return lp


def logpost_literal_sp(u, d, jacobian=False):
    """Series-Parallel[_pos]_modelcode.txt, line by line.  u: torch float64 tensor; d: dict from
    oracle.model_sp.prep_series_parallel.  Returns -inf when the validity check ``real<lower=0> x_sum_raw`` fails."""
    Ks, Kp, Nf = d['Ks'], d['Kp'], d['Nf']
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    As, Ap, Z, freq = t(d['As']), t(d['Ap']), t(d['Z']), t(d['freq'])
    L0s, L1s, L2s = (t(x) for x in d['Ls'])
    L0p, L1p, L2p = (t(x) for x in d['Lp'])
    # transformed data (:26-31)
    Rinf_vec = torch.cat((torch.ones(Nf, dtype=torch.float64), torch.zeros(Nf, dtype=torch.float64)))
    induc_vec = torch.cat((torch.zeros(Nf, dtype=torch.float64), 2 * math.pi * freq))
    pos = 0
    logjac = torch.zeros((), dtype=torch.float64)

    def take(n, lower0):
        nonlocal pos, logjac
        raw = u[pos:pos + n]
        pos += n
        if lower0:
            logjac = logjac + raw.sum()
            return torch.exp(raw)
        return raw

    # parameters (:32-49)
    Rinf_raw = take(1, True)[0]
    induc_raw = take(1, True)[0]
    xs = take(Ks, d['pos'])
    xp_raw = take(Kp, True)
    sigma_res_raw = take(1, True)[0]
    alpha_prop_raw = take(1, True)[0]
    alpha_re_raw = take(1, True)[0]
    alpha_im_raw = take(1, True)[0]
    ups_s_raw = take(Ks, True)
    ups_p_raw = take(Kp, True)
    d0s, d1s, d2s = take(1, True)[0], take(1, True)[0], take(1, True)[0]
    d0p, d1p, d2p = take(1, True)[0], take(1, True)[0], take(1, True)[0]
    assert pos == u.numel()
    # transformed parameters (:50-85)
    Rinf = Rinf_raw * 100
    induc = induc_raw * d['induc_scale']
    xp = xp_raw * d['xp_scale']
    qs = torch.sqrt(d0s * (L0s @ xs) ** 2 + d1s * (L1s @ xs) ** 2 + d2s * (L2s @ xs) ** 2)
    qp = torch.sqrt(d0p * (L0p @ xp_raw) ** 2 + d1p * (L1p @ xp_raw) ** 2 + d2p * (L2p @ xp_raw) ** 2)
    x_sum_raw = xs.sum() + xp_raw.sum()
    if x_sum_raw.item() < 0:
        return torch.tensor(-float('inf'), dtype=torch.float64)
    x_sum = x_sum_raw * d['x_sum_invscale']
    sigma_res = sigma_res_raw * 0.05
    alpha_prop = alpha_prop_raw * 0.05
    alpha_re = alpha_re_raw * 0.05
    alpha_im = alpha_im_raw * 0.05
    Y_hat = Ap @ xp
    Y_hat_re, Y_hat_im = Y_hat[:Nf], Y_hat[Nf:]
    Z_hat_p = torch.cat((Y_hat_re / (Y_hat_re ** 2 + Y_hat_im ** 2), -Y_hat_im / (Y_hat_re ** 2 + Y_hat_im ** 2)))
    Z_hat = Z_hat_p + As @ xs + Rinf * Rinf_vec + induc * induc_vec
    Z_hat_re = torch.cat((Z_hat[:Nf], Z_hat[:Nf]))
    Z_hat_im = torch.cat((Z_hat[Nf:], Z_hat[Nf:]))
    sigma_tot = torch.sqrt(d['sigma_min'] ** 2 + sigma_res ** 2 + (alpha_prop * Z_hat) ** 2
                           + (alpha_re * Z_hat_re) ** 2 + (alpha_im * Z_hat_im) ** 2)
    ups_s = ups_s_raw * 0.15
    ups_p = ups_p_raw * 0.15
    dups_s = 0.5 * (ups_s[1:-1] - 0.5 * (ups_s[:-2] + ups_s[2:])) / ups_s[1:-1]
    dups_p = 0.5 * (ups_p[1:-1] - 0.5 * (ups_p[:-2] + ups_p[2:])) / ups_p[1:-1]
    # model (:86-107)
    lp = sum(_inv_gamma_lpdf(x, 5.0, 5.0) for x in (d0s, d1s, d2s, d0p, d1p, d2p))
    lp = lp + _std_normal_lpdf(x_sum)
    lp = lp + _inv_gamma_lpdf(ups_s_raw, d['ups_alpha'], d['ups_beta']) \
        + _inv_gamma_lpdf(ups_p_raw, d['ups_alpha'], d['ups_beta'])
    lp = lp + _std_normal_lpdf(Rinf_raw) + _std_normal_lpdf(induc_raw)
    lp = lp + _normal_lpdf(qs, 0.0, ups_s) + _normal_lpdf(qp, 0.0, ups_p)
    lp = lp + _std_normal_lpdf(dups_s) + _std_normal_lpdf(dups_p)
    lp = lp + _normal_lpdf(Z, Z_hat, sigma_tot)
    lp = lp + _std_normal_lpdf(sigma_res_raw) + _std_normal_lpdf(alpha_prop_raw) + _std_normal_lpdf(alpha_re_raw) \
        + _std_normal_lpdf(alpha_im_raw)
    if jacobian:
        lp = lp + logjac
    return lp


def logpost_literal_s2p(u, d, jacobian=False):
    """Series-2Parallel[_pos]_modelcode.txt, line by line (parameters :39-61, transformed parameters :62-103,
    model :104-129).  d: dict from oracle.model_sp.prep_series_2parallel."""
    Ks, Kp1, Kp2, Nf = d['Ks'], d['Kp'], d['Kp2'], d['Nf']
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    As, Ap1, Ap2, Z, freq = t(d['As']), t(d['Ap']), t(d['Ap2']), t(d['Z']), t(d['freq'])
    L0s, L1s, L2s = (t(x) for x in d['Ls'])
    L0p1, L1p1, L2p1 = (t(x) for x in d['Lp'])
    L0p2, L1p2, L2p2 = (t(x) for x in d['Lp2'])
    Rinf_vec = torch.cat((torch.ones(Nf, dtype=torch.float64), torch.zeros(Nf, dtype=torch.float64)))
    induc_vec = torch.cat((torch.zeros(Nf, dtype=torch.float64), 2 * math.pi * freq))
    pos = 0
    logjac = torch.zeros((), dtype=torch.float64)

    def take(n, lower0):
        nonlocal pos, logjac
        raw = u[pos:pos + n]
        pos += n
        if lower0:
            logjac = logjac + raw.sum()
            return torch.exp(raw)
        return raw

    Rinf_raw = take(1, True)[0]
    induc_raw = take(1, True)[0]
    xs = take(Ks, d['pos'])
    xp1_raw = take(Kp1, True)
    xp2_raw = take(Kp2, True)
    sigma_res_raw, alpha_prop_raw, alpha_re_raw, alpha_im_raw = (take(1, True)[0] for _ in range(4))
    ups_s_raw = take(Ks, True)
    ups_p1_raw = take(Kp1, True)
    ups_p2_raw = take(Kp2, True)
    ds_ = [take(1, True)[0] for _ in range(3)]
    dp1 = [take(1, True)[0] for _ in range(3)]
    dp2 = [take(1, True)[0] for _ in range(3)]
    assert pos == u.numel()
    Rinf = Rinf_raw * 100
    induc = induc_raw * d['induc_scale']
    xp1 = xp1_raw * d['xp_scale']
    xp2 = xp2_raw * d['xp2_scale']
    q = lambda dd, L0, L1, L2, x: torch.sqrt(dd[0] * (L0 @ x) ** 2 + dd[1] * (L1 @ x) ** 2 + dd[2] * (L2 @ x) ** 2)
    qs, qp1, qp2 = q(ds_, L0s, L1s, L2s, xs), q(dp1, L0p1, L1p1, L2p1, xp1_raw), q(dp2, L0p2, L1p2, L2p2, xp2_raw)
    x_sum_raw = xs.sum() + xp1_raw.sum() + xp2_raw.sum()
    if x_sum_raw.item() < 0:
        return torch.tensor(-float('inf'), dtype=torch.float64)
    x_sum = x_sum_raw * d['x_sum_invscale']
    sigma_res, alpha_prop, alpha_re, alpha_im = (0.05 * v for v in (sigma_res_raw, alpha_prop_raw, alpha_re_raw,
                                                                    alpha_im_raw))

    def zp(Y):
        re, im = Y[:Nf], Y[Nf:]
        return torch.cat((re / (re ** 2 + im ** 2), -im / (re ** 2 + im ** 2)))
    Z_hat = zp(Ap1 @ xp1) + zp(Ap2 @ xp2) + As @ xs + Rinf * Rinf_vec + induc * induc_vec
    Z_hat_re = torch.cat((Z_hat[:Nf], Z_hat[:Nf]))
    Z_hat_im = torch.cat((Z_hat[Nf:], Z_hat[Nf:]))
    sigma_tot = torch.sqrt(d['sigma_min'] ** 2 + sigma_res ** 2 + (alpha_prop * Z_hat) ** 2
                           + (alpha_re * Z_hat_re) ** 2 + (alpha_im * Z_hat_im) ** 2)
    dups = lambda ur: (lambda ups: 0.5 * (ups[1:-1] - 0.5 * (ups[:-2] + ups[2:])) / ups[1:-1])(ur * 0.15)
    lp = sum(_inv_gamma_lpdf(x, 5.0, 5.0) for x in ds_ + dp1 + dp2)
    lp = lp + _std_normal_lpdf(x_sum)
    for ur in (ups_s_raw, ups_p1_raw, ups_p2_raw):
        lp = lp + _inv_gamma_lpdf(ur, d['ups_alpha'], d['ups_beta']) + _std_normal_lpdf(dups(ur))
    lp = lp + _std_normal_lpdf(Rinf_raw) + _std_normal_lpdf(induc_raw)
    lp = lp + _normal_lpdf(qs, 0.0, ups_s_raw * 0.15) + _normal_lpdf(qp1, 0.0, ups_p1_raw * 0.15) \
        + _normal_lpdf(qp2, 0.0, ups_p2_raw * 0.15)
    lp = lp + _normal_lpdf(Z, Z_hat, sigma_tot)
    lp = lp + _std_normal_lpdf(sigma_res_raw) + _std_normal_lpdf(alpha_prop_raw) + _std_normal_lpdf(alpha_re_raw) \
        + _std_normal_lpdf(alpha_im_raw)
    if jacobian:
        lp = lp + logjac
    return lp
