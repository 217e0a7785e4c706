"""GPU: fused log-posterior + analytic-gradient kernel (bdrt_logpost_grad) against the oracle's restatement of the Stan
programs, for every model variant, both modes, with and without the Jacobian.
Tolerances: |d lp| <= 1e-11 |lp|, max|d grad| <= 1e-9 max|grad| (FP64; the DMMA products sum in a different order)."""
import numpy as np
import pytest
import torch

from helpers import gpu_problem, load_spectrum, oracle_batch
from oracle import model as omod

pytestmark = pytest.mark.gpu


def _check(ds, prob, u, spec, jac):
    lp, grad = prob.logpost_grad(torch.tensor(u), spec=spec, jacobian=jac)
    lp, grad = lp.cpu().numpy(), grad.cpu().numpy()
    for c in range(u.shape[0]):
        b = spec[c] if spec is not None else c % len(ds)
        lo, go = omod.logpost(u[c], ds[b], jacobian=jac)
        assert abs(lp[c] - lo) <= 1e-11 * abs(lo), (c, lp[c], lo)
        assert np.max(np.abs(grad[c] - go)) <= 1e-9 * np.max(np.abs(go)), (c, np.max(np.abs(grad[c] - go)))


@pytest.mark.parametrize('nonneg', [False, True])
@pytest.mark.parametrize('outliers', [False, True])
@pytest.mark.parametrize('mode', ['optimize', 'sample'])
def test_logpost_S_shape(nonneg, outliers, mode, resident_A):
    """data/simulated shape: Nf = 81, default basis K = 101 (D = 211 / 373)."""
    names = ['ZARC_uniform_0.25', 'ZARC-RL_uniform_0.25', '2ZARC_uniform_0.25']
    freq = load_spectrum(names[0])[0]
    ds = oracle_batch(freq, [load_spectrum(n)[1] for n in names], mode=mode, nonneg=nonneg, outliers=outliers)
    prob = gpu_problem(ds)
    assert prob.D == omod.n_params(ds[0])
    rng = np.random.RandomState(5)
    u = rng.uniform(-2, 2, (11, prob.D))  # 11 columns: one full group of 8 slots + a ragged one
    spec = rng.randint(0, 3, 11)
    for jac in (False, True):
        _check(ds, prob, u, spec, jac)


def test_logpost_B_shape_many_columns(resident_A):
    """benchmark shape: Nf = 70, K = 100, shared grid; more columns than one wave of CTAs."""
    from bayes_drt_b200 import synth
    freq, Z, _ = synth.make_spectra(40, seed=11)
    _, bf = synth.bench_grid()
    ds = oracle_batch(freq.numpy(), list(Z.numpy()), basis_freq=bf.numpy(), mode='optimize')
    prob = gpu_problem(ds)
    assert prob.D == 209
    rng = np.random.RandomState(6)
    u = rng.uniform(-2, 2, (1500, prob.D))
    lp, grad = prob.logpost_grad(torch.tensor(u))
    lp, grad = lp.cpu().numpy(), grad.cpu().numpy()
    for c in list(range(0, 1500, 97)) + [1499]:
        lo, go = omod.logpost(u[c], ds[c % 40])
        assert abs(lp[c] - lo) <= 1e-11 * abs(lo)
        assert np.max(np.abs(grad[c] - go)) <= 1e-9 * np.max(np.abs(go))


def test_logpost_near_optimum_and_nonfinite(resident_A):
    """points near a MAP optimum (boundary parameters at exp(-20)), and a point that overflows -> non-finite lp, not a
    crash (Stan rejects such points)."""
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    ds = oracle_batch(freq, [Z], mode='optimize')
    prob = gpu_problem(ds)
    rng = np.random.RandomState(7)
    u = rng.uniform(-1, 1, (3, prob.D))
    u[0, 1] = -20.0
    u[0, 3 + ds[0]['K']] = -30.0
    u[1, 2:2 + ds[0]['K']] *= 0.01
    _check(ds, prob, u, None, False)
    u[2, -1] = 800.0  # d2_strength = exp(800) = inf
    lp, grad = prob.logpost_grad(torch.tensor(u))
    assert not np.isfinite(lp[2].item()) or not np.all(np.isfinite(grad[2].cpu().numpy()))


def test_logpost_non_toeplitz_grid():
    """measurement frequencies off the basis grid (jittered): A is not Toeplitz, the engine must keep the dense A."""
    rng = np.random.RandomState(8)
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    fj = freq * np.exp(rng.uniform(-0.05, 0.05, len(freq)))
    ds = oracle_batch(fj, [Z, Z[::-1].copy()], basis_freq=np.logspace(6.5, -2.5, 91), mode='sample', nonneg=True)
    prob = gpu_problem(ds)
    u = rng.uniform(-1.5, 1.5, (9, prob.D))
    for jac in (False, True):
        _check(ds, prob, u, None, jac)
