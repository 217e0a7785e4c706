"""CPU emulation of the control flow of csrc/newton.cu (damped Newton polish of MAP estimates) on the oracle model.

Why it exists: the benchmark-shape MAP parity test (tests/test_gpu_map_benchmark.py) showed (i) 13 % of the spectra
running all 200 Newton iterations without reaching max|g| < 1e-9 and (ii) 2 of 256 spectra ending 4e-4 .. 2e-3 of the
peak away from the oracle's optimum.  Both were reproduced here, without a GPU, by running the kernel's logic
statement by statement on oracle.model.logpost, and the fixes were tried here first:

  * the rounding floor of max|g| is 1e-10 .. 1e-8 depending on the spectrum: below it steps are still accepted but the
    norm only wanders -> stop after `stall` iterations that did not halve the best value; accept a step on a halved
    gradient norm when the Armijo test is decided by the noise of f (as oracle/newton.py);
  * a central-difference Hessian (`central=True`) changes nothing but the cost;
  * lower=0 coordinates: `jump='size'` is the first version's rule (Newton step < -0.5 and theta < exp(-6) -> u = -40,
    frozen, released on g < -gtol, which a gradient scaled by theta = 4e-18 never meets): it freezes the scaled
    inductance of spectrum 115 at 0 although its optimum is exp(-21), err 4.46e-4.  `jump='sign'` (adopted) keeps the
    rule but checks the sign of d lp / d theta at the floor on the next iteration and sends a wrongly frozen coordinate
    back for good.  `jump='unit'` (two consecutive steps of about -1) was tried in between: it never freezes the error
    scales, whose signature is -1/2 (they enter squared), and leaves sigma_res ~ 4e-5 instead of 0.

    python scripts/proto_newton_logic.py 115,240,30,0,1,2 [size|unit|sign]

prints, per spectrum and for three random starts (oracle L-BFGS first): oracle polish vs emulation -- max|g|,
iterations, releases, f, and the largest relative difference of x, R_inf, sigma_res."""
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np  # noqa: E402

from bayes_drt_b200 import synth  # noqa: E402
from helpers import oracle_batch  # noqa: E402
from oracle import lbfgs as olb, model as omod, newton as onew  # noqa: E402

N = 256  # the batch of tests/test_gpu_map_benchmark.py
BS = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else [115, 240, 30]
JUMP = sys.argv[2] if len(sys.argv) > 2 else 'sign'
freq, Z, _ = synth.make_spectra(N, seed=20240601)
_, bf = synth.bench_grid()
DS = oracle_batch(freq.numpy(), list(Z.numpy()), basis_freq=bf.numpy(), mode='optimize')


def fg(d, u):
    lp, g = omod.logpost(u, d)
    return -lp, -g


def cuda_newton(d, u0, jump='sign', central=False, halfrule=True, stall=6, tries_max=12, max_iter=200, gtol=1e-9,
                h0=1e-6):
    u = u0.copy()
    D = len(u)
    xs = omod.param_slices(d)['x']
    expc = np.array([not (xs.start <= i < xs.stop) for i in range(D)])  # Series, nonneg=False: only x is unconstrained
    state = np.zeros(D, int)  # 0 free, 1 frozen at the floor, 2 free for good
    cnt = np.zeros(D, int)
    ujump = np.zeros(D)
    f, g = fg(d, u)
    mu, it, best, nst, nrel = 1e-6, 0, np.inf, 0, 0
    while True:
        rel = (state == 1) & ((g < 0) if jump != 'size' else (g < -gtol))
        if rel.any():
            if jump == 'size':
                state[rel] = 0
            else:
                state[rel] = 2
                u[rel] = ujump[rel]
                f, g = fg(d, u)
            nrel += int(rel.sum())
        frozen = state == 1
        gmax = np.max(np.abs(g[~frozen]))
        if gmax < gtol or it >= max_iter:
            break
        if stall:
            if gmax < 0.5 * best:
                best, nst = gmax, 0
            elif gmax < 1e-6:
                nst += 1
                if nst >= stall:
                    break
        it += 1
        H = np.zeros((D, D))
        for j in np.where(~frozen)[0]:
            h = h0 * max(1.0, abs(u[j]))
            up = u.copy()
            up[j] = u[j] + h
            if central:
                um = u.copy()
                um[j] = u[j] - h
                H[j] = (fg(d, up)[1] - fg(d, um)[1]) / ((up[j] - u[j]) + (u[j] - um[j]))
            else:
                H[j] = (fg(d, up)[1] - g) / (up[j] - u[j])
        accepted = False
        for _ in range(tries_max):
            Hs = 0.5 * (H + H.T)
            Hs[np.diag_indices(D)] += mu * np.maximum(np.abs(np.diag(Hs)), 1e-12)
            Hs[frozen, :] = 0
            Hs[:, frozen] = 0
            Hs[frozen, frozen] = 1
            try:
                Lc = np.linalg.cholesky(Hs)
            except np.linalg.LinAlgError:
                mu *= 10
                continue
            step = np.linalg.solve(Lc.T, np.linalg.solve(Lc, np.where(frozen, 0.0, -g)))
            unit = (step > -1.5) & (step < -0.7)
            if jump == 'unit':
                jmp = expc & (state == 0) & unit & (cnt >= 1) & (u + step < -6.0)
            else:
                jmp = expc & (state == 0) & (step < -0.5) & (u + step < -6.0)
            utry = np.where(jmp, -40.0, u + step)
            gd = g @ (utry - u)
            with np.errstate(all='ignore'):
                ftry, gtry = fg(d, utry)
            free = ~frozen & ~jmp
            gtmax = np.max(np.abs(gtry[free])) if np.isfinite(ftry) else np.inf
            if np.isfinite(ftry) and (ftry <= f + 1e-4 * gd + 1e-13 * abs(f) or
                                      (halfrule and gmax < 1e-5 and gtmax < 0.5 * gmax and ftry < f + 1e-9 * abs(f))):
                ujump[jmp] = u[jmp]
                u, g, f = utry, gtry, ftry
                state[jmp] = 1
                mu = max(mu * 0.1, 1e-12)
                cnt = np.where(unit, cnt + 1, 0)
                accepted = True
                break
            mu *= 10
            if mu > 1e12:
                break
        if not accepted:
            break
    return dict(u=u, gnorm=gmax, iters=it, nrel=nrel)


def run(b):
    d = DS[b]
    sl = omod.param_slices(d)['x']
    out = []

    def func(u):
        lp, g = omod.logpost(u, d)
        if not np.isfinite(lp) or not np.all(np.isfinite(g)):
            return None
        return -lp, -g
    for seed in (100 + b, 200 + b, 300 + b):
        u0 = np.random.RandomState(seed).uniform(-2, 2, omod.n_params(d))
        with np.errstate(all='ignore'):
            ue = olb.minimize(func, u0, max_iter=50000)['x']
            o = onew.polish(func, ue, max_iter=120)
            c = cuda_newton(d, ue, jump=JUMP)
        co, cc = omod.constrain(o['x'], d), omod.constrain(c['u'], d)
        err = max(np.max(np.abs(c['u'][sl] - o['x'][sl])) / np.max(np.abs(o['x'][sl])),
                  abs(co['sigma_res'] - cc['sigma_res']) / (co['sigma_res'] + 0.2),
                  abs(co['Rinf'] - cc['Rinf']) / (co['Rinf'] + 0.2))
        out.append((seed, o['gnorm'], o['iters'], o['f'], c['gnorm'], c['iters'], c['nrel'], func(c['u'])[0], err))
    return b, out


if __name__ == '__main__':
    with mp.get_context('fork').Pool(min(os.cpu_count() or 1, 16)) as pool:
        for b, out in pool.imap_unordered(run, BS):
            for o in out:
                print(f'b={b} seed {o[0]} oracle max|g| {o[1]:.2e} it {o[2]} f {o[3]:.10f} | emulation ({JUMP}) max|g| '
                      f'{o[4]:.2e} it {o[5]} released {o[6]} f {o[7]:.10f} | err {o[8]:.2e}', flush=True)
