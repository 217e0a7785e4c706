"""Thin tensor-level wrappers over the C ABI (one Python function per extern "C" entry point of include/bdrt.h).

All tensors are CUDA float64; every function enqueues work on torch's current stream.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import BdrtError, LbfgsOpts, NewtonOpts, NutsOpts, RidgeOpts, SeriesData, SeriesInfo, context, f64, ptr


def build_A(freq, tau, epsilon, kernel='DRT', dist_type='series', symmetry='planar', bc='transmissive', ct=False,
            k_ct=None, device=None):
    """A_re, A_im of matrices.construct_A for one grid (freq [Nf]) or a batch of grids (freq [G, Nf]).
    tau: [K] shared or [G, K].  Returns ([G,] Nf, K) tensors."""
    ctx = context(device)
    freq = f64(freq, ctx.device)
    tau = f64(tau, ctx.device)
    single = freq.dim() == 1
    fr = freq.reshape(1, -1) if single else freq
    G, Nf = fr.shape
    tau_per_grid = tau.dim() == 2
    if tau_per_grid and tau.shape[0] != G:
        raise ValueError('tau batch dimension must match freq')
    K = tau.shape[-1]
    if ct and k_ct is None:
        raise ValueError('k_ct must be supplied if ct==True')  # matrices.py:42-43
    A_re = torch.empty((G, Nf, K), dtype=torch.float64, device=ctx.device)
    A_im = torch.empty_like(A_re)
    rc = ctx.lib.bdrt_build_A(ctx._h, ptr(fr), G, Nf, ptr(tau), K, int(tau_per_grid), C.c_double(float(epsilon)),
                              _lib.KERNEL[kernel], _lib.DIST[dist_type], _lib.SYM[symmetry or 'planar'],
                              _lib.BC[bc or 'transmissive'], int(bool(ct)), C.c_double(float(k_ct or 0.0)),
                              ptr(A_re), ptr(A_im))
    ctx.check(rc)
    return (A_re[0], A_im[0]) if single else (A_re, A_im)


def build_L(freq, tau, epsilon, order, device=None):
    ctx = context(device)
    freq = f64(freq, ctx.device)
    tau = f64(tau, ctx.device)
    single = freq.dim() == 1
    fr = freq.reshape(1, -1) if single else freq
    G, N = fr.shape
    K = tau.shape[-1]
    L = torch.empty((G, N, K), dtype=torch.float64, device=ctx.device)
    rc = ctx.lib.bdrt_build_L(ctx._h, ptr(fr), G, N, ptr(tau), K, int(tau.dim() == 2), C.c_double(float(epsilon)),
                              int(order), ptr(L))
    ctx.check(rc)
    return L[0] if single else L


def build_M(freq, epsilon, order, toeplitz, device=None):
    ctx = context(device)
    freq = f64(freq, ctx.device)
    single = freq.dim() == 1
    fr = freq.reshape(1, -1) if single else freq
    G, K = fr.shape
    M = torch.empty((G, K, K), dtype=torch.float64, device=ctx.device)
    rc = ctx.lib.bdrt_build_M(ctx._h, ptr(fr), G, K, C.c_double(float(epsilon)), int(order), int(bool(toeplitz)),
                              ptr(M))
    ctx.check(rc)
    return M[0] if single else M


class SeriesProblem:
    """Owns the device tensors behind one bdrt_series_data (keeps them alive for the duration of the calls)."""

    def __init__(self, A, Z, freq, L, nonneg=False, outliers=False, sigma_min=0.002, ups_alpha=0.05, ups_beta=0.1,
                 induc_scale=1.0, sigma_out_lambda=10.0, sigma_out_alpha=2.0, sigma_out_beta=1.0, device=None,
                 Ap=None, Lp=None, x_sum_invscale=0.0, xp_scale=1.0, parallel=False, Ap2=None, Lp2=None, xp2_scale=1.0):
        """Series family: A, L describe the single DRT.  Series-Parallel (pass Ap, Lp): A, L describe the series
        distribution, Ap [2Nf, Kp] / Lp [3, Kp, Kp] the parallel one (Stan data of inversion.py:1886-1959).
        ``parallel=True``: the single distribution is a parallel one (Stan program 'Parallel', x is lower=0).
        Series-2Parallel: additionally pass Ap2 / Lp2 / xp2_scale (second parallel distribution, sorted-name order)."""
        self.ctx = context(device)
        dev = self.ctx.device
        self.A = f64(A, dev)
        self.Z = f64(Z, dev)
        self.freq = f64(freq, dev)
        self.L = f64(L, dev)
        if self.Z.dim() != 2:
            raise ValueError('Z must be [B, 2Nf]')
        self.B, n2 = self.Z.shape
        self.Nf = n2 // 2
        self.K = self.A.shape[-1]
        self.per_spectrum_grid = self.A.dim() == 3
        if self.A.shape[-2] != n2 or self.freq.shape[-1] != self.Nf or tuple(self.L.shape) != (3, self.K, self.K):
            raise ValueError('inconsistent shapes in SeriesProblem')
        self.series_parallel = Ap is not None
        d = SeriesData()
        if parallel and (self.series_parallel or outliers):
            raise ValueError("parallel=True is the single-distribution 'Parallel' program (no Ap, no outlier model)")
        self.parallel = bool(parallel)
        nonneg = bool(nonneg) or self.parallel
        self.two_parallel = Ap2 is not None
        if self.two_parallel and not self.series_parallel:
            raise ValueError('Ap2 needs Ap')
        d.model = (_lib.MODEL_SERIES_2PARALLEL if self.two_parallel else
                   (_lib.MODEL_SERIES_PARALLEL if self.series_parallel else
                    (_lib.MODEL_PARALLEL if parallel else _lib.MODEL_SERIES))) \
            | (_lib.MODEL_POS if nonneg else 0) | (_lib.MODEL_OUTLIERS if outliers else 0)
        d.Nf, d.K, d.B = self.Nf, self.K, self.B
        d.per_spectrum_grid = int(self.per_spectrum_grid)
        d.A, d.Z, d.freq, d.L = self.A.data_ptr(), self.Z.data_ptr(), self.freq.data_ptr(), self.L.data_ptr()
        d.sigma_min, d.ups_alpha, d.ups_beta, d.induc_scale = sigma_min, ups_alpha, ups_beta, induc_scale
        d.sigma_out_lambda, d.sigma_out_alpha, d.sigma_out_beta = sigma_out_lambda, sigma_out_alpha, sigma_out_beta
        self.Kp = 0
        if self.series_parallel:
            self.Ap, self.Lp = f64(Ap, dev), f64(Lp, dev)
            self.Kp = self.Ap.shape[-1]
            if self.Ap.shape[-2] != n2 or (self.Ap.dim() == 3) != self.per_spectrum_grid \
                    or tuple(self.Lp.shape) != (3, self.Kp, self.Kp):
                raise ValueError('inconsistent shapes of the parallel distribution in SeriesProblem')
            d.Kp, d.Ap, d.Lp = self.Kp, self.Ap.data_ptr(), self.Lp.data_ptr()
            d.x_sum_invscale, d.xp_scale = float(x_sum_invscale), float(xp_scale)
        self.Kp2 = 0
        if self.two_parallel:
            self.Ap2, self.Lp2 = f64(Ap2, dev), f64(Lp2, dev)
            self.Kp2 = self.Ap2.shape[-1]
            if self.Ap2.shape[-2] != n2 or (self.Ap2.dim() == 3) != self.per_spectrum_grid \
                    or tuple(self.Lp2.shape) != (3, self.Kp2, self.Kp2):
                raise ValueError('inconsistent shapes of the second parallel distribution in SeriesProblem')
            d.Kp2, d.Ap2, d.Lp2, d.xp2_scale = self.Kp2, self.Ap2.data_ptr(), self.Lp2.data_ptr(), float(xp2_scale)
        self.c = d
        # structure of the matrices (Toeplitz / band flags, taps), found once: the solver calls then only enqueue work
        self.info = SeriesInfo()
        if self.B > 0:
            self.ctx.check(self.ctx.lib.bdrt_series_analyze(self.ctx._h, C.byref(d), C.byref(self.info)))
            d.info = C.pointer(self.info)
        self.nonneg, self.outliers = bool(nonneg), bool(outliers)
        self.D = int(self.ctx.lib.bdrt_num_params(C.byref(d)))
        self.P = int(self.ctx.lib.bdrt_num_outputs(C.byref(d)))

    def subset(self, a, b):
        """Spectra a .. b - 1 of this problem as a problem of their own: a view (same device tensors, same analysis of the
        matrices), used to run a large batch through the sampler in bounded pieces."""
        import copy
        sub = copy.copy(self)
        d = SeriesData()
        C.memmove(C.byref(d), C.byref(self.c), C.sizeof(SeriesData))
        n2 = 2 * self.Nf
        sub.Z = self.Z[a:b]
        d.B, d.Z = b - a, sub.Z.data_ptr()
        if self.per_spectrum_grid:
            sub.A, sub.freq = self.A[a:b], self.freq[a:b]
            d.A, d.freq = sub.A.data_ptr(), sub.freq.data_ptr()
            if self.series_parallel:
                sub.Ap = self.Ap[a:b]
                d.Ap = sub.Ap.data_ptr()
            if self.two_parallel:
                sub.Ap2 = self.Ap2[a:b]
                d.Ap2 = sub.Ap2.data_ptr()
        sub.c, sub.B = d, b - a
        return sub

    # -- log_prob / grad_log_prob test hook
    def logpost_grad(self, u, spec=None, jacobian=False):
        u = f64(u, self.ctx.device)
        n = u.shape[0]
        lp = torch.empty(n, dtype=torch.float64, device=u.device)
        grad = torch.empty_like(u)
        sp = None if spec is None else torch.as_tensor(spec, dtype=torch.int32, device=u.device).contiguous()
        self.ctx.check(self.ctx.lib.bdrt_logpost_grad(self.ctx._h, C.byref(self.c), ptr(u), ptr(sp), n,
                                                      int(bool(jacobian)), ptr(lp), ptr(grad)))
        return lp, grad

    def map_lbfgs(self, u0, max_iter=2000, history=5, **tol):
        u = f64(u0, self.ctx.device).clone()
        if tuple(u.shape) != (self.B, self.D):
            raise ValueError(f'u0 must be [{self.B}, {self.D}]')
        o = LbfgsOpts()
        self.ctx.lib.bdrt_lbfgs_default_opts(C.byref(o))
        o.max_iter, o.history = int(max_iter), int(history)
        for k, v in tol.items():
            setattr(o, k, v)
        lp = torch.empty(self.B, dtype=torch.float64, device=u.device)
        iters = torch.empty(self.B, dtype=torch.int32, device=u.device)
        nev = torch.empty_like(iters)
        status = torch.empty_like(iters)
        self.ctx.check(self.ctx.lib.bdrt_map_lbfgs(self.ctx._h, C.byref(self.c), C.byref(o), ptr(u), ptr(lp),
                                                   ptr(iters), ptr(nev), ptr(status)))
        return dict(u=u, lp=lp, iters=iters, n_eval=nev, status=status)

    def map_newton(self, u0, max_iter=200, gtol=1e-9, fd_step=1e-6):
        u = f64(u0, self.ctx.device).clone()
        o = NewtonOpts()
        self.ctx.lib.bdrt_newton_default_opts(C.byref(o))
        o.max_iter, o.gtol, o.fd_step = int(max_iter), float(gtol), float(fd_step)
        lp = torch.empty(self.B, dtype=torch.float64, device=u.device)
        gnorm = torch.empty_like(lp)
        iters = torch.empty(self.B, dtype=torch.int32, device=u.device)
        nev = torch.empty_like(iters)
        self.ctx.check(self.ctx.lib.bdrt_map_newton(self.ctx._h, C.byref(self.c), C.byref(o), ptr(u), ptr(lp),
                                                    ptr(gnorm), ptr(iters), ptr(nev)))
        return dict(u=u, lp=lp, gnorm=gnorm, iters=iters, n_eval=nev)

    def nuts(self, u0, chains=2, warmup=200, samples=200, seed=1234, adapt_delta=0.9, adapt_t0=10.0,
             max_treedepth=10, spectrum_offset=0, keep_draws=True, spectrum_ids=None):
        u0 = f64(u0, self.ctx.device)
        if tuple(u0.shape) != (self.B, chains, self.D):
            raise ValueError(f'u0 must be [{self.B}, {chains}, {self.D}]')
        o = NutsOpts()
        self.ctx.lib.bdrt_nuts_default_opts(C.byref(o))
        o.chains, o.warmup, o.samples, o.max_treedepth = int(chains), int(warmup), int(samples), int(max_treedepth)
        o.adapt_delta, o.adapt_t0, o.seed, o.spectrum_offset = float(adapt_delta), float(adapt_t0), int(seed), \
            int(spectrum_offset)
        dev = u0.device
        ids = None
        if spectrum_ids is not None:
            ids = torch.as_tensor(spectrum_ids, dtype=torch.int64, device=dev).contiguous()
            if ids.numel() != self.B:
                raise ValueError('spectrum_ids must have one entry per spectrum')
            o.spectrum_ids = ids.data_ptr()
        draws = torch.empty((self.B, chains, samples, self.D), dtype=torch.float64, device=dev) if keep_draws else None
        stepsize = torch.empty((self.B, chains), dtype=torch.float64, device=dev)
        nleap = torch.empty((self.B, chains), dtype=torch.int64, device=dev)
        ndiv = torch.empty((self.B, chains), dtype=torch.int32, device=dev)
        nmax = torch.empty_like(ndiv)
        acc = torch.empty_like(stepsize)
        self.ctx.check(self.ctx.lib.bdrt_nuts(self.ctx._h, C.byref(self.c), C.byref(o), ptr(u0), ptr(draws),
                                              ptr(stepsize), ptr(nleap), ptr(ndiv), ptr(nmax), ptr(acc)))
        return dict(draws=draws, stepsize=stepsize, n_leapfrog=nleap, n_divergent=ndiv, n_maxdepth=nmax, accept=acc)

    def constrain(self, u, spec=None):
        u = f64(u, self.ctx.device)
        lead = u.shape[:-1]
        u2 = u.reshape(-1, self.D)
        n = u2.shape[0]
        out = torch.empty((n, self.P), dtype=torch.float64, device=u.device)
        sp = None if spec is None else torch.as_tensor(spec, dtype=torch.int32, device=u.device).contiguous()
        self.ctx.check(self.ctx.lib.bdrt_constrain(self.ctx._h, C.byref(self.c), ptr(u2), ptr(sp), n, ptr(out)))
        return out.reshape(*lead, self.P)

    def split_outputs(self, out):
        """Named views of bdrt_constrain's packed output."""
        K, Nf = self.K + self.Kp + self.Kp2, self.Nf
        d = {'x': out[..., :K], 'Rinf': out[..., K], 'induc': out[..., K + 1], 'sigma_res': out[..., K + 2],
             'alpha_prop': out[..., K + 3], 'alpha_re': out[..., K + 4], 'alpha_im': out[..., K + 5],
             'sigma_tot': out[..., K + 6:K + 6 + 2 * Nf]}
        if self.outliers:
            d['sigma_out'] = out[..., K + 6 + 2 * Nf:]
        if self.series_parallel:
            d['xs'], d['xp'] = out[..., :self.K], out[..., self.K:self.K + self.Kp]
        if self.two_parallel:
            d['xp1'], d['xp2'] = d['xp'], out[..., self.K + self.Kp:K]
        return d


def summarize(draws, percentiles=(), want_mean=True, device=None):
    """Posterior mean and percentiles (np.percentile semantics, linear interpolation) of draws [G, S, P] over the S axis.
    Returns (mean [G, P] or None, quant [nq, G, P] or None)."""
    ctx = context(device)
    draws = f64(draws, ctx.device)
    if draws.dim() != 3:
        raise ValueError('draws must be [G, S, P]')
    G, S, P = draws.shape
    probs = (C.c_double * max(len(percentiles), 1))(*[float(p) / 100.0 for p in percentiles])
    mean = torch.empty((G, P), dtype=torch.float64, device=ctx.device) if want_mean else None
    quant = torch.empty((len(percentiles), G, P), dtype=torch.float64, device=ctx.device) if len(percentiles) else None
    ctx.check(ctx.lib.bdrt_summarize(ctx._h, ptr(draws), G, S, P, probs, len(percentiles), ptr(mean), ptr(quant)))
    return mean, quant


SUMMARIZE_MAX_DRAWS = 16384  # merged draws per parameter the shared-memory sort of bdrt_summarize holds on B200


def diagnostics(draws, chains, device=None):
    """Split R-hat and rank-normalised bulk ESS of draws [G, chains * n, P] (chain-major) -> (rhat [G, P], ess [G, P])."""
    ctx = context(device)
    draws = f64(draws, ctx.device)
    if draws.dim() != 3 or draws.shape[1] % chains:
        raise ValueError('draws must be [G, chains * n, P]')
    G, S, P = draws.shape
    rhat = torch.empty((G, P), dtype=torch.float64, device=ctx.device)
    ess = torch.empty((G, P), dtype=torch.float64, device=ctx.device)
    ctx.check(ctx.lib.bdrt_diagnostics(ctx._h, ptr(draws), G, int(chains), S // chains, P, ptr(rhat), ptr(ess)))
    return rhat, ess


def qp_bound(P, q, lb, device=None):
    ctx = context(device)
    P, q, lb = f64(P, ctx.device), f64(q, ctx.device), f64(lb, ctx.device)
    B, n = q.shape
    x = torch.empty_like(q)
    kkt = torch.empty(B, dtype=torch.float64, device=ctx.device)
    iters = torch.empty(B, dtype=torch.int32, device=ctx.device)
    ctx.check(ctx.lib.bdrt_qp_bound(ctx._h, ptr(P), ptr(q), ptr(lb), B, n, ptr(x), ptr(kkt), ptr(iters)))
    return x, kkt, iters


def ridge_fit(WA_re, WA_im, WZ_re, WZ_im, Pen, Lmat, penalty='discrete', nonneg=True, max_iter=20, xtol=1e-3,
              hl_beta=2.5, lambda_0=1e-2, reg_ord=(0.0, 0.0, 1.0), L1_penalty=0.0, epsilon=1.0, fit_inductance=True,
              hl_fbeta=None, stop_rule=1, device=None):
    """stop_rule: see bdrt_ridge_opts in include/bdrt.h (0 numpy 0/0 = NaN semantics, 1 unchanged-on-the-bound counts as
    converged).  Returns coef, lam, iters (hyper-iterations), converged, n_factor (Cholesky factorisations)."""
    ctx = context(device)
    dev = ctx.device
    WA_re, WA_im = f64(WA_re, dev), f64(WA_im, dev)
    WZ_re, WZ_im = f64(WZ_re, dev), f64(WZ_im, dev)
    Pen = f64(Pen, dev)
    Lm = None if Lmat is None else f64(Lmat, dev)
    B, Nf = WZ_re.shape
    n = WA_re.shape[-1]
    o = RidgeOpts()
    ctx.lib.bdrt_ridge_default_opts(C.byref(o))
    o.penalty = {'discrete': 0, 'integral': 1}[penalty]
    o.nonneg, o.max_iter, o.xtol, o.hl_beta, o.lambda_0 = int(nonneg), int(max_iter), xtol, hl_beta, lambda_0
    for i in range(3):
        o.reg_ord[i] = float(reg_ord[i])
    o.L1_penalty, o.epsilon, o.fit_inductance = float(L1_penalty), float(epsilon), int(fit_inductance)
    o.hl_fbeta = 0.0 if hl_fbeta is None else float(hl_fbeta)
    o.stop_rule = int(stop_rule)
    coef = torch.empty((B, n), dtype=torch.float64, device=dev)
    lam = torch.empty((B, 3, n), dtype=torch.float64, device=dev)
    iters = torch.empty(B, dtype=torch.int32, device=dev)
    conv = torch.empty(B, dtype=torch.int32, device=dev)
    nfac = torch.empty(B, dtype=torch.int32, device=dev)
    ctx.check(ctx.lib.bdrt_ridge_fit(ctx._h, C.byref(o), ptr(WA_re), ptr(WA_im), int(WA_re.dim() == 3), ptr(WZ_re),
                                     ptr(WZ_im), ptr(Pen), ptr(Lm), B, Nf, n - 2, ptr(coef), ptr(lam), ptr(iters),
                                     ptr(conv), ptr(nfac)))
    return dict(coef=coef, lam=lam, iters=iters, converged=conv, n_factor=nfac)


def peak_fp64(device=None):
    ctx = context(device)
    a, b = C.c_double(), C.c_double()
    ctx.check(ctx.lib.bdrt_peak_fp64(ctx._h, C.byref(a), C.byref(b)))
    return a.value, b.value
