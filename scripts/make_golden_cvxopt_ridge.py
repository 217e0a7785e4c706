"""Golden fixture from REAL cvxopt output: hyper-parametric ridge fits the reference's authors saved for their paper.

``code_EchemActa/comparisons/hyper-ridge/results/obj_<circuit>_<noise>_fbeta=<v>.pkl`` are objects saved by an earlier
version of the reference's ``ridge_fit`` (``hl_fbeta`` hyper-prior, ``penalty='discrete'``, ``reg_ord=2``,
``scale_Z=False``, ``dZ=False``, ``max_iter=50``; notebook ``hyper-ridge run fits.ipynb``).  Their ``_iter_history`` keeps,
for EVERY hyper-iteration, the lambda vectors, the coefficient vector ``cvxopt.solvers.qp`` returned, cvxopt's primal
objective and its duality gap.  cvxopt is not installable here, so these are the only QP solutions in the tree that
cvxopt itself computed.  This script reads them in place (build container only), works out which quadratic program each
stored solution belongs to, checks that reading against the stored objective, and writes
``tests/golden/cvxopt_ridge.npz``.

How that version set up the QP (found by reproducing the stored objectives, to 1e-13 relative; asserted below at 1e-11):
  min 1/2 c'Pc + q'c, c >= 0;  c = [R_inf, L * 1e4, x_1 .. x_K]
  P = A_re'A_re + A_im'A_im + Lam^1/2 (L2'L2) Lam^1/2,   q = -A_re'Z' + A_im'Z''
  A_im is stored for -Z'' (the sign of today's matrices.construct_A(.., 'imag') flipped), its inductance column enters
  the solve scaled by 1e-4, Lam = the lambda vector of the same history entry, and the next entry's lambda vector is
  lambda_0 / ((L2 c)^2 / (max (L2 c)^2 * hl_fbeta) + 1) -- today's _hyper_lambda_fbeta (inversion.py:956-964).
Only fits with unity weights (the ``uniform`` noise model) are used: the other weighting schemes of that version are not
recoverable from the objects.

  <name>/freq, Z, tau, eps, fbeta, lambda_0
  <name>/proj/*          seeded random projections of the stored A_re, A_im, L2 (the tests rebuild them with the oracle)
  <name>/P0_triu, q      A_re'A_re + A_im'A_im (upper triangle) and q of that run; P = P0 + Lam^1/2 (L2'L2) Lam^1/2
  <name>/lam [n, K+2]    lambda vector of the QP of hyper-iteration i = 0 .. n-1
  <name>/coef [n, K+2]   cvxopt's solution, <name>/cost [n] its primal objective, <name>/gap [n] its duality gap
  <name>/lam_next [n, K+2]  the lambda vector the reference computed from that solution (NaN for the last one)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from make_golden_stan_map import load_obj, proj_vectors  # noqa: E402

REF = '/root/reference'
RES = os.path.join(REF, 'code_EchemActa/comparisons/hyper-ridge/results')
INDUC_SCALE = 1e-4
PICK = [('ZARC', '1.0', '1'), ('ZARC', '0.25', '10'), ('2RC', '1.0', '1'), ('2RC', '2.5', '0.1'), ('Gerischer', '1.0', '1'),
        ('Gerischer', '0.25', '100'), ('2ZARC', '1.0', '10'), ('RC', '2.5', '1'), ('RC', '0.25', '0.01')]


def system(d, Z, lam):
    A_re, A_im, L2 = d['A_re'], d['A_im'].copy(), d['L2']
    A_im[:, 1] *= INDUC_SCALE
    P = A_re.T @ A_re + A_im.T @ A_im + np.diag(lam ** 0.5) @ (L2.T @ L2) @ np.diag(lam ** 0.5)
    q = -A_re.T @ Z.real + A_im.T @ Z.imag
    return P, q


def main():
    import pandas as pd
    out, names = {}, []
    for circ, noise, fb in PICK:
        name = f'{circ}_uniform_{noise}_fbeta={fb}'
        path = os.path.join(RES, f'obj_{name}.pkl')
        if not os.path.exists(path):
            print('missing', name)
            continue
        d = load_obj(path)
        df = pd.read_csv(os.path.join(REF, 'data/simulated', f'Z_{circ}_uniform_{noise}.csv'))
        idx = np.argsort(df['Freq'].values)[::-1]
        freq = df['Freq'].values[idx]
        Z = (df['Zreal'].values + 1j * df['Zimag'].values)[idx]
        assert np.allclose(freq, d['f_train'], rtol=1e-12) and d['_Z_scale'] == 1
        h = d['_iter_history']
        lam, coef, cost, gap, nxt = [], [], [], [], []
        for i in range(len(h)):
            c = np.asarray(h[i]['coef'], dtype=np.float64)
            lv = np.asarray(h[i]['lambda_vectors'][2], dtype=np.float64)
            P, q = system(d, Z, lv)
            f = 0.5 * c @ P @ c + q @ c
            assert abs(f - h[i]['cost']) <= 1e-11 * abs(f), (name, i, f, h[i]['cost'])  # the reading of the QP is right
            lam.append(lv)
            coef.append(c)
            cost.append(float(h[i]['cost']))
            gap.append(float(h[i]['result']['gap']))
            # the lambda vector the reference computed from this solution is the one of the next entry
            nxt.append(np.asarray(h[i + 1]['lambda_vectors'][2], dtype=np.float64) if i + 1 < len(h) else np.full(len(lv), np.nan))
        p = name + '/'
        out[p + 'freq'], out[p + 'Z'], out[p + 'tau'] = freq, Z, np.asarray(d['tau'], dtype=np.float64)
        out[p + 'eps'], out[p + 'fbeta'] = np.array(float(d['_epsilon'])), np.array(float(fb))
        out[p + 'lambda_0'] = np.array(float(h[0]['lambda_vectors'][0][0]))
        for mn in ('A_re', 'A_im', 'L2'):
            M = np.asarray(d[mn], dtype=np.float64)
            l, r = proj_vectors(*M.shape)
            out[p + f'proj/{mn}/r'], out[p + f'proj/{mn}/l'] = M @ r, l @ M
        # the lambda-independent part of the system exactly as that run had it (its imaginary kernel matrix comes from an
        # older quadrature, 1e-7 off today's: too coarse to compare objectives at cvxopt's gap of 1e-7 relative)
        P0, q = system(d, Z, np.zeros(len(lam[0])))
        out[p + 'P0_triu'], out[p + 'q'] = P0[np.triu_indices(len(q))], q
        out[p + 'lam'], out[p + 'coef'] = np.array(lam), np.array(coef)
        out[p + 'cost'], out[p + 'gap'], out[p + 'lam_next'] = np.array(cost), np.array(gap), np.array(nxt)
        names.append(name)
        print(f'{name}: {len(lam)} QPs, K = {len(d["tau"])}, gaps {min(gap):.1e} .. {max(gap):.1e}')
    out['names'] = np.array(names)
    dst = os.path.join(ROOT, 'tests', 'golden', 'cvxopt_ridge.npz')
    np.savez_compressed(dst, **out)
    print(f'wrote {dst}: {len(names)} fits, {os.path.getsize(dst) / 1024:.0f} kB')


if __name__ == '__main__':
    main()
