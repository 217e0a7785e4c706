"""Generate tests/golden/matrices.npz by running the REFERENCE's own matrices.py.

Run in the build container only (needs /root/reference; the GPU box has no
reference tree):

    python scripts/make_golden_matrices.py

The reference module is imported verbatim (bayes_drt/matrices.py: construct_A
:120, construct_L :268, construct_M :366); nothing is copied into this repo.
Each case stores its inputs next to the reference outputs so the tests need no
reference at run time.
"""
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, '/root/reference')
warnings.simplefilter('ignore')
from bayes_drt import matrices as ref  # noqa: E402

out = {}


def default_tau(freq):
    # inversion.py:2191-2199
    tmin = np.log10(1 / (2 * np.pi * np.max(freq))) - 1
    tmax = np.log10(1 / (2 * np.pi * np.min(freq))) + 1
    return np.logspace(tmin, tmax, int(10 * (tmax - tmin) + 1))


def add_case(name, freq, tau, eps, kernel='DRT', dist_type='series', symmetry='planar', bc='', ct=False, k_ct=None,
             with_LM=False):
    kw = dict(tau=tau, basis='gaussian', epsilon=eps, kernel=kernel, dist_type=dist_type, symmetry=symmetry, bc=bc,
              ct=ct, k_ct=k_ct)
    out[f'{name}/freq'] = freq
    out[f'{name}/tau'] = tau
    out[f'{name}/eps'] = np.float64(eps)
    out[f'{name}/meta'] = np.array([kernel, dist_type, symmetry, bc or '', str(bool(ct)), repr(k_ct)])
    out[f'{name}/A_re'] = ref.construct_A(freq, 'real', **kw)
    out[f'{name}/A_im'] = ref.construct_A(freq, 'imag', **kw)
    if with_LM:
        bf = 1 / (2 * np.pi * tau)
        for o in (0, 1, 2):
            out[f'{name}/L{o}'] = ref.construct_L(bf, tau=tau, basis='gaussian', epsilon=eps, order=o)  # inversion.py:2301-2307
            out[f'{name}/M{o}'] = ref.construct_M(bf, basis='gaussian', order=o, epsilon=eps)  # inversion.py:2296-2299
        # third derivative, fractional and list-mixed orders (matrices.py:278-316)
        for key, o in (('L3', 3), ('Lf0.5', 0.5), ('Lf1.25', 1.25), ('Lmix', [0.2, 0.3, 0.5])):
            out[f'{name}/{key}'] = ref.construct_L(bf, tau=tau, basis='gaussian', epsilon=eps, order=o)
    print(name, out[f'{name}/A_re'].shape)


# S: data/simulated default shape (Nf=81, K=101)
fS = np.logspace(6, -2, 81)
tS = default_tau(fS)
eS = 1 / np.mean(np.diff(np.log(tS)))
add_case('S', fS, tS, eS, with_LM=True)

# B: benchmark shape (Nf=70, K=100), SURVEY section 8d config 4
fB = 10.0 ** (5 - np.arange(70) / 10)
tB = 1 / (2 * np.pi * 10.0 ** (6 - np.arange(100) / 10))
eB = 1 / np.mean(np.diff(np.log(tB)))
add_case('B', fB, tB, eB, with_LM=True)

# RC-ZARC grid of Tutorial 0 (41 freqs), custom basis tau
fR = np.logspace(5, -3, 41)
tR = np.logspace(-2, 3, 51)
add_case('RCZARC', fR, tR, 1 / np.mean(np.diff(np.log(tR))), with_LM=True)

# basis_freq = freq (tau == 1/omega Toeplitz branch, matrices.py:137-141)
fT = np.logspace(4, -1, 26)
add_case('TAUEQ', fT, 1 / (2 * np.pi * fT), 1 / np.mean(np.diff(np.log(1 / fT))))

# jittered, non-Toeplitz grids (full path, matrices.py:243-263), non-default epsilon
rng = np.random.RandomState(7)
fJ = np.sort(10.0 ** (rng.uniform(-2, 5, 23)))[::-1].copy()
tJ = np.sort(10.0 ** (rng.uniform(-7, 2, 29)))
add_case('JIT_eps2', fJ, tJ, 2.0, with_LM=True)
add_case('JIT_eps9', fJ, tJ, 9.0)
add_case('JIT_eps05', fJ, tJ, 0.5)

# DDT kernels (matrices.py:56-112)
fD = np.logspace(6, -2, 41)
tD = 1 / (2 * np.pi * np.logspace(6, -2, 41))
eD = 1 / np.mean(np.diff(np.log(tD)))
add_case('DDT_TP_par', fD, tD, eD, kernel='DDT', dist_type='parallel', symmetry='planar', bc='transmissive')
add_case('DDT_BP_par', fD, tD, eD, kernel='DDT', dist_type='parallel', symmetry='planar', bc='blocking')
add_case('DDT_TP_ser', fD, tD, eD, kernel='DDT', dist_type='series', symmetry='planar', bc='transmissive')
add_case('DDT_BP_ser', fD, tD, eD, kernel='DDT', dist_type='series', symmetry='planar', bc='blocking')
add_case('DDT_BS_par', fD[::2], tD[::2], eD / 2, kernel='DDT', dist_type='parallel', symmetry='spherical',
         bc='blocking')
add_case('DDT_TP_par_ct', fD[::2], tD[::2], eD / 2, kernel='DDT', dist_type='parallel', symmetry='planar',
         bc='transmissive', ct=True, k_ct=3.0)
add_case('DDT_BP_ser_ct', fD[::2], tD[::2], eD / 2, kernel='DDT', dist_type='series', symmetry='planar',
         bc='blocking', ct=True, k_ct=0.5)
# DDT on a jittered grid (full path)
add_case('DDT_TP_par_JIT', fJ, tJ, 3.0, kernel='DDT', dist_type='parallel', symmetry='planar', bc='transmissive')

dst = os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'matrices.npz')
np.savez_compressed(dst, **out)
print('wrote', dst, os.path.getsize(dst) / 1e3, 'kB')
