"""GPU: batched on-device L-BFGS MAP driver against the oracle's restatement of Stan's L-BFGS from identical inits."""
import numpy as np
import pytest
import torch

from helpers import gpu_problem, load_spectrum, oracle_batch
from oracle import lbfgs as olb, model as omod

pytestmark = pytest.mark.gpu


def _func(d):
    def f(u):
        lp, g = omod.logpost(u, d)
        if not np.isfinite(lp) or not np.all(np.isfinite(g)):
            return None
        return -lp, -g
    return f


def test_lbfgs_first_iterations_match_oracle(resident_A):
    """Same algorithm, same init: after a fixed small number of iterations (before round-off differences between the
    DMMA products and numpy's BLAS have been amplified by the ill-conditioning) the iterates agree closely."""
    freq, Z = load_spectrum('ZARC_uniform_0.25')
    ds = oracle_batch(freq, [Z, load_spectrum('2ZARC_uniform_0.25')[1]], mode='optimize')
    prob = gpu_problem(ds)
    rng = np.random.RandomState(0)
    u0 = rng.uniform(-2, 2, (2, prob.D))
    with np.errstate(all='ignore'):
        # Stan's random init U(-2,2) makes the first steepest-descent trial steps land where |lp| ~ 1e266 and whether the
        # *gradient* overflows depends on the order of operations (numpy vs CUDA vs Stan's AD), which changes the
        # line-search path.  Parity of the algorithm is therefore checked from a sane start: 40 oracle iterations in.
        u0 = np.stack([olb.minimize(_func(ds[b]), u0[b], max_iter=40)['x'] for b in range(2)])
        for n_it in (1, 5, 25):
            r = prob.map_lbfgs(torch.tensor(u0), max_iter=n_it)
            for b in range(2):
                o = olb.minimize(_func(ds[b]), u0[b], max_iter=n_it)
                assert r['iters'][b].item() == o['iters'] == n_it
                assert r['n_eval'][b].item() == o['n_eval']
                assert abs(r['lp'][b].item() + o['f']) <= 1e-9 * abs(o['f'])
                assert np.max(np.abs(r['u'][b].cpu().numpy() - o['x'])) <= 1e-7 * np.max(np.abs(o['x']))


def test_lbfgs_converges_like_oracle():
    """Run to Stan's own termination: same termination class, objective within 1e-4 relative of the oracle's run (the two
    trajectories decorrelate, so this is a statement about the algorithm, not bit-parity)."""
    freq, Z = load_spectrum('ZARC-RL_uniform_0.25')
    ds = oracle_batch(freq, [Z], mode='optimize')
    prob = gpu_problem(ds)
    rng = np.random.RandomState(1)
    u0 = rng.uniform(-2, 2, (1, prob.D))
    r = prob.map_lbfgs(torch.tensor(u0), max_iter=50000)
    with np.errstate(all='ignore'):
        o = olb.minimize(_func(ds[0]), u0[0], max_iter=50000)
    assert r['status'][0].item() in (21, 31) and o['code'] in (21, 31)
    assert abs(r['lp'][0].item() + o['f']) <= 1e-3 * abs(o['f'])
    lo, go = omod.logpost(r['u'][0].cpu().numpy(), ds[0])
    assert abs(lo - r['lp'][0].item()) <= 1e-10 * abs(lo)


def test_lbfgs_batch_independent_of_batching(resident_A):
    """Results per spectrum are bitwise independent of which other spectra share the CTA / the batch."""
    from bayes_drt_b200 import synth
    freq, Z, _ = synth.make_spectra(24, seed=2)
    _, bf = synth.bench_grid()
    ds = oracle_batch(freq.numpy(), list(Z.numpy()), basis_freq=bf.numpy(), mode='optimize')
    prob = gpu_problem(ds)
    rng = np.random.RandomState(3)
    u0 = torch.tensor(rng.uniform(-2, 2, (24, prob.D)))
    r_all = prob.map_lbfgs(u0, max_iter=200)
    prob2 = gpu_problem(ds[5:8])
    r_sub = prob2.map_lbfgs(u0[5:8], max_iter=200)
    assert torch.equal(r_all['u'][5:8], r_sub['u'])
    assert torch.equal(r_all['lp'][5:8], r_sub['lp'])
    assert torch.all(r_all['lp'] > prob.logpost_grad(u0)[0])


def test_map_polish_reaches_the_oracle_optimum():
    """North star: MAP DRT coefficients within 1e-5 (relative to the largest coefficient) -- evaluated, as SURVEY section 7
    prescribes, against the oracle's tightly converged optimum (damped Newton to max|grad| < 1e-9), which is unique for
    the default model; both sides start from the same L-BFGS result.  R_inf and the error-model scales likewise;
    boundary parameters (inductance, alpha_*: theta -> 0) are compared absolutely."""
    from oracle import newton as onew
    names = ['ZARC_uniform_0.25', '2ZARC_uniform_0.25']
    freq = load_spectrum(names[0])[0]
    ds = oracle_batch(freq, [load_spectrum(n)[1] for n in names], mode='optimize')
    prob = gpu_problem(ds)
    rng = np.random.RandomState(0)
    u0 = torch.tensor(rng.uniform(-2, 2, (2, prob.D)))
    r = prob.map_lbfgs(u0, max_iter=50000)
    p = prob.map_newton(r['u'])
    assert (p['gnorm'] < 1e-7).all(), p['gnorm']
    assert (p['lp'] >= r['lp'] - 1e-9).all()
    out = prob.split_outputs(prob.constrain(p['u']))
    for b in range(2):
        with np.errstate(all='ignore'):
            o = onew.polish(_func(ds[b]), r['u'][b].cpu().numpy(), max_iter=80)
        assert o['gnorm'] < 1e-8
        c = omod.constrain(o['x'], ds[b])
        x = out['x'][b].cpu().numpy()
        assert abs(p['lp'][b].item() + o['f']) <= 1e-10 * abs(o['f'])
        assert np.max(np.abs(x - c['x'])) <= 1e-5 * np.max(np.abs(c['x'])), np.max(np.abs(x - c['x'])) / np.max(np.abs(c['x']))
        assert abs(out['Rinf'][b].item() - c['Rinf']) <= 1e-5 * c['Rinf']
        assert abs(out['sigma_res'][b].item() - c['sigma_res']) <= 1e-5 * c['sigma_res'] + 1e-9
        for nm in ('induc', 'alpha_prop', 'alpha_re', 'alpha_im'):
            assert abs(out[nm][b].item() - c[nm]) <= 1e-5 * abs(c[nm]) + 1e-7, nm
