"""CPU: the oracle's matrices against the reference's own matrices.py outputs (tests/golden/matrices.npz, generated
by scripts/make_golden_matrices.py which imports the reference verbatim)."""
import os

import numpy as np
import pytest

from oracle import matrices as om

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'matrices.npz'))
CASES = sorted(set(k.split('/')[0] for k in G.files))


def case_kwargs(n):
    kern, dt, sym, bc, ct, kct = G[n + '/meta']
    return dict(kernel=str(kern), dist_type=str(dt), symmetry=str(sym), bc=str(bc) or None, ct=(ct == 'True'),
                k_ct=eval(str(kct)))


def a_tolerance(n, R_re, R_im):
    """DRT: 1e-10 relative per entry (north star).  DDT: the reference's own small entries carry cancellation noise of
    ~1e-14 x the largest complex entry (SURVEY 8c 'DDT caveat'), so: |dA| <= 1e-10 |A| + 1e-13 max|A_re + i A_im|."""
    if n.startswith('DDT'):
        return 1e-10, 1e-13 * np.nanmax(np.hypot(R_re, R_im))
    return 1e-10, 0.0


@pytest.mark.parametrize('name', CASES)
def test_construct_A(name):
    f, t, e = G[name + '/freq'], G[name + '/tau'], float(G[name + '/eps'])
    R = {p: G[f'{name}/A_{p}'] for p in ('re', 'im')}
    rtol, atol = a_tolerance(name, R['re'], R['im'])
    for part, key in (('real', 're'), ('imag', 'im')):
        A = om.construct_A(f, part, tau=t, epsilon=e, **case_kwargs(name))
        ok = np.isfinite(R[key])  # the reference itself returns NaN for spherical DDT at tiny x (0/0)
        assert ok.mean() > 0.9
        err = np.abs(A - R[key])[ok]
        assert np.all(err <= rtol * np.abs(R[key][ok]) + atol + 1e-300), (name, part, err.max())


@pytest.mark.parametrize('name', [c for c in CASES if c + '/L0' in G.files])
def test_construct_L_M(name):
    t, e = G[name + '/tau'], float(G[name + '/eps'])
    bf = 1 / (2 * np.pi * t)
    for o in (0, 1, 2):
        L = om.construct_L(bf, tau=t, epsilon=e, order=o)
        M = om.construct_M(bf, order=o, epsilon=e)
        RL, RM = G[f'{name}/L{o}'], G[f'{name}/M{o}']
        assert np.max(np.abs(L - RL)) <= 1e-12 * np.abs(RL).max()
        assert np.max(np.abs(M - RM)) <= 1e-12 * np.abs(RM).max()


L_ORDERS = (('L3', 3), ('Lf0.5', 0.5), ('Lf1.25', 1.25), ('Lmix', [0.2, 0.3, 0.5]))


@pytest.mark.parametrize('name', [c for c in CASES if c + '/L3' in G.files])
def test_construct_L_fractional_and_mixed_orders(name):
    """third derivative, fractional and list-mixed orders (matrices.py:278-316)"""
    t, e = G[name + '/tau'], float(G[name + '/eps'])
    bf = 1 / (2 * np.pi * t)
    for key, o in L_ORDERS:
        L = om.construct_L(bf, tau=t, epsilon=e, order=o)
        RL = G[f'{name}/{key}']
        assert np.max(np.abs(L - RL)) <= 1e-12 * np.abs(RL).max(), key
