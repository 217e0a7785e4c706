"""Oracle restatement of the reference's kernel / penalty matrices (numpy, FP64).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ``bayes_drt/matrices.py``:
  * basis function           matrices.py:8-24   (gaussian only, as Inverter documents)
  * integrands               matrices.py:27-117
  * construct_A (trapz)      matrices.py:120-265  -- default ``integrate_method='trapz'``:
                             ``np.trapz`` over ``np.linspace(-20, 20, 1000)`` (:235-238, :261-263)
  * construct_L              matrices.py:268-325
  * construct_M              matrices.py:328-411

The reference's Toeplitz shortcut (matrices.py:145-242) only changes *which*
entries are integrated, not their values, so the oracle always integrates every
entry.  Pinned against the reference itself through tests/golden/matrices.npz
(made by scripts/make_golden_matrices.py, which imports the reference verbatim).
"""
import numpy as np

_Y = np.linspace(-20.0, 20.0, 1000)  # matrices.py:236 / :262


def _trapz(f, y):
    # np.trapz semantics: sum(diff(y) * (f[1:] + f[:-1]) / 2) along the last axis
    d = np.diff(y)
    return np.sum(d * (f[..., 1:] + f[..., :-1]) * 0.5, axis=-1)


def _tanh_c(x):
    """Overflow-safe complex tanh for Re(x) >= 0."""
    e = np.exp(-2.0 * x)
    return (1.0 - e) / (1.0 + e)


def _ZD(y, w, t, symmetry, bc, ct, k_ct):
    """Diffusion impedance Z_D(y; w_n, t_m)  (matrices.py:56-94)."""
    if ct:
        x = np.sqrt(t * np.exp(y) * (k_ct + 1j * w))
    else:
        x = np.sqrt(1j * w * t * np.exp(y))
    th = _tanh_c(x)
    if bc == 'blocking':
        if symmetry == 'planar':
            return 1.0 / (th * x)
        elif symmetry == 'spherical':
            return th / (x - th)
        raise ValueError(f'Invalid symmetry {symmetry}')
    elif bc == 'transmissive':
        if symmetry == 'planar':
            return th / x
        raise ValueError(f'Invalid symmetry {symmetry}')
    raise ValueError(f'Invalid bc {bc}')


def integrand(y, w, t, part, epsilon, kernel='DRT', dist_type='series', symmetry='planar', bc=None,
              ct=False, k_ct=None):
    """g(y; w_n, t_m, eps), broadcasting over all arguments (matrices.py:45-112)."""
    phi = np.exp(-(epsilon * y) ** 2)  # matrices.py:11-13
    if kernel == 'DRT':
        if dist_type != 'series':
            raise ValueError('dist_type for DRT kernel must be series')
        e2 = np.exp(2.0 * (y + np.log(w * t)))
        if part == 'real':
            return phi / (1.0 + e2)  # matrices.py:48-49
        return -phi * np.exp(y) * w * t / (1.0 + e2)  # matrices.py:51-52
    elif kernel == 'DDT':
        zd = _ZD(y, w, t, symmetry, bc, ct, k_ct)
        val = 1.0 / zd if dist_type == 'parallel' else zd  # matrices.py:97-110
        return phi * (val.real if part == 'real' else val.imag)
    raise ValueError(f'Invalid kernel {kernel}')


def construct_A(frequencies, part, tau=None, epsilon=1.0, kernel='DRT', dist_type='series',
                symmetry='planar', bc=None, ct=False, k_ct=None):
    """A[n, m] = trapz_y g(y; 2*pi*f_n, tau_m)   (matrices.py:120-265, trapz path)."""
    frequencies = np.asarray(frequencies, dtype=np.float64)
    omega = 2.0 * np.pi * frequencies
    tau = 1.0 / omega if tau is None else np.asarray(tau, dtype=np.float64)
    A = np.empty((len(omega), len(tau)))
    with np.errstate(over='ignore', invalid='ignore'):
        for n, w in enumerate(omega):
            f = integrand(_Y[None, :], w, tau[:, None], part, epsilon, kernel, dist_type, symmetry, bc, ct, k_ct)
            A[n] = _trapz(f, _Y)
    return A


def construct_L(frequencies, tau=None, epsilon=1.0, order=1):
    """L[n, m] = d^order/dy^order exp(-(eps*y)^2) at y = ln(1/(w_n tau_m))  (matrices.py:268-325)."""
    omega = 2.0 * np.pi * np.asarray(frequencies, dtype=np.float64)
    tau = 1.0 / omega if tau is None else np.asarray(tau, dtype=np.float64)
    y = np.log(1.0 / (omega[:, None] * tau[None, :]))
    e = np.exp(-(epsilon * y) ** 2)
    d0 = e
    d1 = -2 * epsilon ** 2 * y * e
    d2 = (-2 * epsilon ** 2 + 4 * epsilon ** 4 * y ** 2) * e
    d3 = (12 * epsilon ** 4 * y - 8 * epsilon ** 6 * y ** 3) * e
    if isinstance(order, list):
        f0, f1, f2 = order
        return f0 * d0 + f1 * d1 + f2 * d2
    if order == 0:
        return d0
    if order == 1:
        return d1
    if order == 2:
        return d2
    if order == 3:
        return d3
    if 0 < order < 1:
        return (1 - order) * d0 + order * d1
    if 1 < order < 2:
        return (2 - order) * d1 + (order - 1) * d2
    raise ValueError('Order must be between 0 and 3')


def is_loguniform(frequencies):
    """utils.py:134-140 verbatim semantics -- NOTE the quirk: std/mean of *negative* log-steps is <= 0.01 for ANY
    descending array, so every descending grid counts as log-uniform."""
    fdiff = np.diff(np.log(frequencies))
    return bool(np.std(fdiff) / np.mean(fdiff) <= 0.01)


def construct_M(frequencies, order=1, epsilon=1.0):
    """M[n, m] = integral of products of basis-function derivatives  (matrices.py:328-411).

    Mirrors the reference's symmetric-Toeplitz shortcut (matrices.py:396-405): when ``is_loguniform(frequencies)``
    the matrix is ``toeplitz(first column)`` -- which, through the quirk above, happens for every descending
    ``frequencies`` (always the case inside Inverter, inversion.py:2296-2299), log-uniform or not."""
    frequencies = np.asarray(frequencies, dtype=np.float64)
    omega = 2.0 * np.pi * frequencies
    if is_loguniform(frequencies):
        full = construct_M_full(frequencies, order, epsilon)
        c = full[:, 0]
        idx = np.abs(np.arange(len(c))[:, None] - np.arange(len(c))[None, :])
        return c[idx]
    return construct_M_full(frequencies, order, epsilon)


def construct_M_full(frequencies, order=1, epsilon=1.0):
    """All entries from the closed form (matrices.py:340-360, :406-409)."""
    omega = 2.0 * np.pi * np.asarray(frequencies, dtype=np.float64)
    a = epsilon * np.log(1.0 / (omega[:, None] * (1.0 / omega)[None, :]))
    e = np.exp(-(a ** 2 / 2))
    c = (np.pi / 2) ** 0.5
    m0 = c / epsilon * e
    m1 = -c * epsilon * (-1 + a ** 2) * e
    m2 = c * epsilon ** 3 * (3 - 6 * a ** 2 + a ** 4) * e
    if isinstance(order, list):
        f0, f1, f2 = order
        return f0 * m0 + f1 * m1 + f2 * m2
    return {0: m0, 1: m1, 2: m2}[order]
