"""Host-side mirror of the reference's ``bayes_drt/matrices.py`` interface, computed by the sm_100a kernels.

Same names, argument meaning and error behaviour as the reference functions (matrices.py: construct_A :120,
construct_L :268, construct_M :366); inputs may be numpy arrays or torch tensors, outputs are CUDA float64 tensors.
``integrate_method='quad'`` is not offered: the reference's ``Inverter`` never uses it (inversion.py:2250-2269 take
the default 'trapz', which is also what the CUDA kernel reproduces node for node).
"""
import numpy as np
import torch

from . import capi


def rel_round(x, precision):
    """utils.py:113-131: round to `precision` significant digits."""
    x = np.asarray(x, dtype=np.float64)
    x_scale = np.floor(np.log10(x + 1e-30))
    digits = (precision - x_scale).astype(int)
    return np.array([round(float(xi), int(di)) for xi, di in zip(np.atleast_1d(x), np.atleast_1d(digits))])


def is_loguniform(frequencies):
    """utils.py:134-140, including its quirk (True for any descending array)."""
    f = np.asarray(torch.as_tensor(frequencies).cpu(), dtype=np.float64)
    fdiff = np.diff(np.log(f))
    return bool(np.std(fdiff) / np.mean(fdiff) <= 0.01)


def construct_A(frequencies, part, tau=None, basis='gaussian', fit_inductance=False, epsilon=1, kernel='DRT',
                dist_type='series', symmetry='planar', bc=None, ct=False, k_ct=None, integrate_method='trapz'):
    if basis != 'gaussian':
        raise ValueError(f'Invalid basis {basis}. Options are gaussian')  # matrices.py:22-23
    if integrate_method != 'trapz':
        raise NotImplementedError("integrate_method='quad' is not implemented (the reference default 'trapz' is)")
    if part not in ('real', 'imag'):
        raise ValueError(f'Invalid part {part}')
    if kernel not in ('DRT', 'DDT'):
        raise ValueError(f'Invalid kernel {kernel}. Options are DRT and DDT')  # matrices.py:114-115
    if kernel == 'DRT' and dist_type != 'series':
        raise ValueError('dist_type for DRT kernel must be series')  # matrices.py:53-54
    if kernel == 'DDT':
        if dist_type not in ('series', 'parallel'):
            raise ValueError(f'Invalid dist_type {dist_type}. Options are series and parallel')  # matrices.py:111-112
        if bc == 'blocking' and symmetry not in ('planar', 'spherical'):
            raise ValueError(f'Invalid symmetry {symmetry}. Options are planar or spherical for bc=blocking')
        if bc == 'transmissive' and symmetry != 'planar':
            raise ValueError(f'Invalid symmetry {symmetry}. Symmetry must be planar for bc=transmissive')
        if bc not in ('blocking', 'transmissive'):
            raise ValueError(f'Invalid bc {bc}')
    f = torch.as_tensor(frequencies, dtype=torch.float64)
    t = 1.0 / (2 * np.pi * f) if tau is None else torch.as_tensor(tau, dtype=torch.float64)
    A_re, A_im = capi.build_A(f, t, epsilon, kernel=kernel, dist_type=dist_type, symmetry=symmetry or 'planar',
                              bc=bc or 'transmissive', ct=ct, k_ct=k_ct)
    return A_re if part == 'real' else A_im


def construct_L(frequencies, tau=None, basis='gaussian', epsilon=1, order=1):
    if basis != 'gaussian':
        raise ValueError(f'Invalid basis {basis}. Options are gaussian')
    f = torch.as_tensor(frequencies, dtype=torch.float64)
    t = 1.0 / (2 * np.pi * f) if tau is None else torch.as_tensor(tau, dtype=torch.float64)
    # matrices.py:278-316: a list [f0, f1, f2] mixes the 0th/1st/2nd derivative, a fractional order interpolates
    # linearly between its two neighbouring integer orders; the integer orders come from the kernel
    if isinstance(order, list):
        f0, f1, f2 = order
        mix = ((0, f0), (1, f1), (2, f2))
    elif order in (0, 1, 2, 3):
        return capi.build_L(f, t, epsilon, int(order))
    elif 0 < order < 1:
        mix = ((0, 1 - order), (1, order))
    elif 1 < order < 2:
        mix = ((1, 2 - order), (2, order - 1))
    else:
        raise ValueError('Order must be between 0 and 3')  # matrices.py:315-316
    L = None
    for o, w in mix:
        Lo = capi.build_L(f, t, epsilon, o) * float(w)
        L = Lo if L is None else L + Lo
    return L


def construct_M(frequencies, basis='gaussian', order=1, epsilon=1):
    if basis != 'gaussian':
        raise ValueError(f'Invalid basis {basis}')  # matrices.py:344-345
    if order not in (0, 1, 2):
        raise ValueError(f'Invalid order {order}')  # matrices.py:361-362
    f = torch.as_tensor(frequencies, dtype=torch.float64)
    return capi.build_M(f, epsilon, int(order), toeplitz=is_loguniform(f))
