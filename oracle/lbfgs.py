"""Oracle restatement of Stan 2.19.1's L-BFGS (what ``StanModel.optimizing`` runs, inversion.py:1216).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

[Stan-upstream] -- the optimiser lives in pystan==2.19.1.1 (setup.py:19), absent from the reference tree and from
this image.  Restated from the published algorithm (Stan reference manual "Optimization algorithms"; Nocedal &
Wright alg. 3.5/3.6, 7.4) with Stan's defaults as the reference passes them:
history 5, init_alpha 1e-3, tol_obj 1e-12, tol_rel_obj 1e4, tol_grad 1e-8, tol_rel_grad 1e7, tol_param 1e-8,
Wolfe c1=1e-4 c2=0.9, minAlpha 1e-12, <=20 line-search iterations.  **Path unpinned**: no Stan here to diff
the iterates against (the end points of the paper's saved Stan fits are near-stationary points of this objective,
tests/test_oracle_stan_map.py); the CUDA driver (csrc/lbfgs.cu) implements the same statement and is compared with this
file iterate by iterate.

The objective is f = -log_prob(jacobian=False).
"""
import numpy as np

EPS = np.finfo(np.float64).eps

TERM_SUCCESS, TERM_ABSX, TERM_ABSF, TERM_RELF, TERM_ABSGRAD, TERM_RELGRAD, TERM_MAXIT, TERM_LSFAIL = \
    0, 10, 20, 21, 30, 31, 40, -1


def cubic_interp(df0, x1, f1, df1, lo, hi):
    """Minimiser on [lo, hi] of the cubic through (0, 0, df0) and (x1, f1, df1)."""
    c3 = (-12 * f1 + 6 * x1 * (df0 + df1)) / (x1 * x1 * x1)
    c2 = -(4 * df0 + 2 * df1) / x1 + 6 * f1 / (x1 * x1)
    c1 = df0
    disc = c2 * c2 - 2.0 * c1 * c3
    t_s = np.sqrt(disc) if disc >= 0 else np.nan
    with np.errstate(divide='ignore', invalid='ignore'):
        s1 = -(c2 + t_s) / c3
        s2 = -(c2 - t_s) / c3

    def val(s):
        return s * (s * (s * c3 / 3.0 + c2) / 2.0 + c1)

    min_f, min_x = val(lo), lo
    tmp = val(hi)
    if tmp < min_f:
        min_f, min_x = tmp, hi
    for s in (s1, s2):
        if lo < s < hi:
            tmp = val(s)
            if tmp < min_f:
                min_f, min_x = tmp, s
    return min_x


def _zoom(func, x, f, dfp, c1dfp, c2dfp, p, alo, alo_f, alo_dfp, ahi, ahi_f, ahi_dfp, min_range, counter):
    it = 0
    while True:
        it += 1
        if abs(alo - ahi) < min_range:
            return 1, None
        if it % 5 == 0:
            alpha = 0.5 * (alo + ahi)
        else:
            d1 = alo_dfp + ahi_dfp - 3 * (alo_f - ahi_f) / (alo - ahi)
            rad = d1 * d1 - alo_dfp * ahi_dfp
            d2 = np.sqrt(rad) if rad >= 0 else np.nan
            if ahi < alo:
                d2 = -d2
            with np.errstate(divide='ignore', invalid='ignore'):
                alpha = ahi - (ahi - alo) * (ahi_dfp + d2 - d1) / (ahi_dfp - alo_dfp + 2 * d2)
            lo, hi = min(alo, ahi), max(alo, ahi)
            if (not np.isfinite(alpha)) or alpha < lo + 0.01 * abs(alo - ahi) or alpha > hi - 0.01 * abs(alo - ahi):
                alpha = 0.5 * (alo + ahi)
        while True:
            xn = x + alpha * p
            counter[0] += 1
            res = func(xn)
            if res is not None:
                break
            alpha = 0.5 * (alpha + min(alo, ahi))
            if abs(min(alo, ahi) - alpha) < min_range:
                return 1, None
        fn, gn = res
        new_dfp = gn @ p
        if fn > (f + alpha * c1dfp) or fn >= alo_f:
            ahi, ahi_f, ahi_dfp = alpha, fn, new_dfp
        else:
            if abs(new_dfp) <= -c2dfp:
                return 0, (alpha, xn, fn, gn)
            if new_dfp * (ahi - alo) >= 0:
                ahi, ahi_f, ahi_dfp = alo, alo_f, alo_dfp
            alo, alo_f, alo_dfp = alpha, fn, new_dfp


def wolfe_line_search(func, alpha, p, x0, f0, g0, c1=1e-4, c2=0.9, min_alpha=1e-12, max_its=20, max_restarts=10,
                      counter=None):
    """Bracketing + zoom strong-Wolfe search.  Returns (retcode, (alpha, x1, f1, g1))."""
    dfp = g0 @ p
    c1dfp, c2dfp = c1 * dfp, c2 * dfp
    alpha0, prev_f, prev_dfp = min_alpha, f0, dfp
    nits = restarts = 0
    while True:
        if nits >= max_its:
            return 1, None
        x1 = x0 + alpha * p
        counter[0] += 1
        res = func(x1)
        if res is None:
            if restarts >= max_restarts:
                return 1, None
            alpha = 0.5 * (alpha0 + alpha)
            restarts += 1
            continue
        restarts = 0
        f1, g1 = res
        new_dfp = g1 @ p
        if f1 > f0 + alpha * c1dfp or (f1 >= prev_f and nits > 0):
            return _zoom(func, x0, f0, dfp, c1dfp, c2dfp, p, alpha0, prev_f, prev_dfp, alpha, f1, new_dfp, 1e-16,
                         counter)
        if abs(new_dfp) <= -c2dfp:
            return 0, (alpha, x1, f1, g1)
        if new_dfp >= 0:
            return _zoom(func, x0, f0, dfp, c1dfp, c2dfp, p, alpha, f1, new_dfp, alpha0, prev_f, prev_dfp, 1e-16,
                         counter)
        alpha0, prev_f, prev_dfp = alpha, f1, new_dfp
        alpha *= 10.0
        nits += 1


def minimize(func, x0, max_iter=50000, history=5, init_alpha=1e-3, tol_obj=1e-12, tol_rel_obj=1e4, tol_grad=1e-8,
             tol_rel_grad=1e7, tol_param=1e-8, trace=None):
    """func(x) -> (f, g) or None when f / g is not finite.  Returns dict(x, f, g, iters, n_eval, code)."""
    counter = [1]
    res = func(x0)
    if res is None:
        raise RuntimeError('non-finite objective at the initial point')
    xk, (fk, gk) = np.array(x0, dtype=np.float64), res
    pk = -gk
    S, Y, RHO = [], [], []
    gamma = 1.0
    alpha = init_alpha
    it = 0
    code = TERM_SUCCESS
    xk_1 = fk_1 = gk_1 = pk_1 = None
    while code == TERM_SUCCESS:
        it += 1
        reset = (it == 1)
        while True:
            if reset:
                pk = -gk
            if it > 1 and not reset:
                alpha = min(1.0, 1.01 * cubic_interp(gk_1 @ pk_1, alpha, fk - fk_1, gk @ pk_1, 1e-12, 1.0))
            else:
                alpha = init_alpha
            rc, out = wolfe_line_search(func, alpha, pk, xk, fk, gk, counter=counter)
            if rc:
                if reset:
                    return dict(x=xk, f=fk, g=gk, iters=it, n_eval=counter[0], code=TERM_LSFAIL)
                reset = True
                continue
            break
        alpha, xn, fn, gn = out
        xk_1, fk_1, gk_1, pk_1 = xk, fk, gk, pk
        xk, fk, gk = xn, fn, gn
        sk, yk = xk - xk_1, gk - gk_1
        grad_norm, step_norm = np.linalg.norm(gk), np.linalg.norm(sk)
        skyk = yk @ sk
        if reset:
            b0 = (yk @ yk) / skyk
            S, Y, RHO = [], [], []
            pk_1 = pk_1 / b0
            alpha = alpha * b0
        gamma = skyk / (yk @ yk)
        S.append(sk)
        Y.append(yk)
        RHO.append(1.0 / skyk)
        if len(S) > history:
            S.pop(0), Y.pop(0), RHO.pop(0)
        # two-loop recursion
        pk = -gk
        al = [0.0] * len(S)
        for i in range(len(S) - 1, -1, -1):
            al[i] = RHO[i] * (S[i] @ pk)
            pk = pk - al[i] * Y[i]
        pk = pk * gamma
        for i in range(len(S)):
            beta = RHO[i] * (Y[i] @ pk)
            pk = pk + (al[i] - beta) * S[i]
        if trace is not None:
            trace.append((it, fk, grad_norm, alpha, counter[0], xk.copy()))
        df = abs(fk_1 - fk)
        if df < tol_obj:
            code = TERM_ABSF
        elif df < tol_rel_obj * max(abs(fk_1), abs(fk), 1.0) * EPS:
            code = TERM_RELF
        elif grad_norm < tol_grad:
            code = TERM_ABSGRAD
        elif -(gk @ pk) / max(abs(fk), 1.0) < tol_rel_grad * EPS:
            code = TERM_RELGRAD
        elif step_norm < tol_param:
            code = TERM_ABSX
        elif it >= max_iter:
            code = TERM_MAXIT
    return dict(x=xk, f=fk, g=gk, iters=it, n_eval=counter[0], code=code)
