"""Damped-Newton polish of a MAP estimate (oracle side).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Not part of the reference: Stan's L-BFGS stops on a relative-objective / relative-gradient test well before the
optimum of this ill-conditioned posterior (cond ~3e7, SURVEY section 7 hard part 1), so "MAP parity" is defined
against the *tightly converged* optimum.  This is the oracle's way to get there: Levenberg-damped Newton on
f = -log_prob(jacobian=False) with a central finite-difference Hessian of the analytic gradient.
The CUDA counterpart is csrc/newton.cu (same differences, damping and acceptance rules; nothing shared).
"""
import numpy as np


def fd_hessian(grad, x, h=1e-6):
    n = len(x)
    H = np.empty((n, n))
    for j in range(n):
        hj = h * max(1.0, abs(x[j]))
        e = np.zeros(n)
        e[j] = hj
        H[:, j] = (grad(x + e) - grad(x - e)) / (2 * hj)
    return 0.5 * (H + H.T)


def polish(func, x0, max_iter=60, gtol=1e-9, verbose=False):
    """func(x) -> (f, g) or None.  Returns dict(x, f, g, iters, gnorm)."""
    x = np.array(x0, dtype=np.float64)
    f, g = func(x)
    mu = 1e-6
    it = 0
    for it in range(1, max_iter + 1):
        gn = np.max(np.abs(g))
        if gn < gtol:
            break
        H = fd_hessian(lambda z: func(z)[1], x)
        dscale = np.maximum(np.abs(np.diag(H)), 1e-12)
        while True:
            try:
                Lc = np.linalg.cholesky(H + mu * np.diag(dscale))
            except np.linalg.LinAlgError:
                mu *= 10
                continue
            step = -np.linalg.solve(Lc.T, np.linalg.solve(Lc, g))
            res = func(x + step)
            # sufficient decrease, or -- at the rounding floor of f (|f| ~ 1e3, decrements ~ gnorm^2), where the Armijo
            # test is decided by noise -- a halved gradient norm: Newton's quadratic phase by its own measure
            if res is not None and (res[0] < f + 1e-4 * (g @ step) or
                                    (gn < 1e-5 and np.max(np.abs(res[1])) < 0.5 * gn and res[0] < f + 1e-9 * abs(f))):
                x = x + step
                f, g = res
                mu = max(mu * 0.1, 1e-12)
                break
            mu *= 10
            if mu > 1e12:
                return dict(x=x, f=f, g=g, iters=it, gnorm=gn, failed=True)
        if verbose:
            print(f'  newton it {it} f {f:.10f} |g|inf {np.max(np.abs(g)):.3e} mu {mu:.1e}')
    return dict(x=x, f=f, g=g, iters=it, gnorm=np.max(np.abs(g)), failed=False)
