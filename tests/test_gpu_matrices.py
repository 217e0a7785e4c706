"""GPU: the sm_100a quadrature kernel (through the C ABI) against the reference's matrices.py outputs
(tests/golden/matrices.npz) at the north star's tolerance -- 1e-10 relative per entry."""
import os

import numpy as np
import pytest
import torch

from test_oracle_matrices import CASES, G, a_tolerance, case_kwargs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', CASES)
def test_build_A_vs_reference(name):
    from bayes_drt_b200 import capi
    f, t, e = G[name + '/freq'], G[name + '/tau'], float(G[name + '/eps'])
    kw = case_kwargs(name)
    A_re, A_im = capi.build_A(torch.tensor(f), torch.tensor(t), e, **kw)
    R = {'re': G[name + '/A_re'], 'im': G[name + '/A_im']}
    rtol, atol = a_tolerance(name, R['re'], R['im'])
    for key, A in (('re', A_re), ('im', A_im)):
        A = A.cpu().numpy()
        ok = np.isfinite(R[key])
        err = np.abs(A - R[key])[ok]
        assert np.all(err <= rtol * np.abs(R[key][ok]) + atol + 1e-300), (name, key, err.max())


@pytest.mark.parametrize('name', [c for c in CASES if c + '/L0' in G.files])
def test_build_L_M_vs_reference(name):
    from bayes_drt_b200 import capi
    from oracle.matrices import is_loguniform
    t, e = G[name + '/tau'], float(G[name + '/eps'])
    bf = 1 / (2 * np.pi * t)
    for o in (0, 1, 2):
        L = capi.build_L(torch.tensor(bf), torch.tensor(t), e, o).cpu().numpy()
        M = capi.build_M(torch.tensor(bf), e, o, toeplitz=is_loguniform(bf)).cpu().numpy()
        RL, RM = G[f'{name}/L{o}'], G[f'{name}/M{o}']
        assert np.max(np.abs(L - RL)) <= 1e-10 * np.abs(RL).max()
        assert np.max(np.abs(M - RM)) <= 1e-10 * np.abs(RM).max()
        # per-entry relative where the entry is not negligible
        big = np.abs(RL) > 1e-12 * np.abs(RL).max()
        assert np.max(np.abs(L - RL)[big] / np.abs(RL)[big]) <= 1e-10


def test_build_A_batched_grids():
    """Per-spectrum grids (config 5): a batch of shifted grids equals one call per grid; ragged last tile sizes."""
    from bayes_drt_b200 import capi
    from oracle import matrices as om
    rng = np.random.RandomState(1)
    Gn, Nf, K = 5, 19, 37
    f = np.stack([10.0 ** (5 - rng.uniform(0, 1) - np.arange(Nf) / 10) for _ in range(Gn)])
    tau = np.stack([1 / (2 * np.pi * 10.0 ** (6 - np.arange(K) / 10 + rng.uniform(0, .3))) for _ in range(Gn)])
    eps = 4.3
    A_re, A_im = capi.build_A(torch.tensor(f), torch.tensor(tau), eps)
    for g in range(Gn):
        for part, A in (('real', A_re), ('imag', A_im)):
            ref = om.construct_A(f[g], part, tau=tau[g], epsilon=eps)
            assert np.max(np.abs(A[g].cpu().numpy() - ref) / np.abs(ref)) <= 1e-10


def test_build_A_errors():
    from bayes_drt_b200 import capi
    from bayes_drt_b200._lib import BdrtError
    f, t = torch.logspace(3, 0, 7, dtype=torch.float64), torch.logspace(-4, 0, 9, dtype=torch.float64)
    with pytest.raises(BdrtError):
        capi.build_A(f, t, 1.0, kernel='DRT', dist_type='parallel')  # matrices.py:53-54
    with pytest.raises(BdrtError):
        capi.build_A(f, t, 1.0, kernel='DDT', dist_type='parallel', symmetry='spherical', bc='transmissive')
    with pytest.raises(ValueError):
        capi.build_A(f, t, 1.0, kernel='DDT', dist_type='parallel', bc='blocking', ct=True)  # matrices.py:42-43
    with pytest.raises(BdrtError):
        capi.build_L(f, t, 1.0, 4)


def test_build_A_more_grids_than_one_launch():
    """> 65535 grids (gridDim.z limit) go in several launches; every grid equals the single-grid result."""
    from bayes_drt_b200 import capi
    G = 70000
    rng = np.random.RandomState(0)
    f = torch.tensor(10 ** (3 - rng.uniform(0, 1, (G, 1)) - np.arange(9)[None, :] / 3))
    tau = torch.tensor(np.logspace(-4, 0, 7))
    A_re, A_im = capi.build_A(f, tau, 1.7)
    assert tuple(A_re.shape) == (G, 9, 7)
    for g in (0, 65534, 65535, 69999):
        a, b = capi.build_A(f[g], tau, 1.7)
        assert torch.equal(a, A_re[g]) and torch.equal(b, A_im[g])
    L = capi.build_L(f, tau, 1.7, 1)
    assert torch.equal(capi.build_L(f[69999], tau, 1.7, 1), L[69999])


@pytest.mark.parametrize('name', [c for c in CASES if c + '/L3' in G.files])
def test_construct_L_fractional_and_mixed_orders(name):
    """matrices.construct_L: third derivative, fractional and list-mixed orders (matrices.py:278-316)"""
    from bayes_drt_b200 import matrices
    from test_oracle_matrices import L_ORDERS
    t, e = G[name + '/tau'], float(G[name + '/eps'])
    bf = 1 / (2 * np.pi * t)
    for key, o in L_ORDERS:
        L = matrices.construct_L(bf, tau=t, epsilon=e, order=o).cpu().numpy()
        RL = G[f'{name}/{key}']
        assert np.max(np.abs(L - RL)) <= 1e-10 * np.abs(RL).max(), key
    with pytest.raises(ValueError, match='Order must be between 0 and 3'):
        matrices.construct_L(bf, tau=t, epsilon=e, order=3.5)
