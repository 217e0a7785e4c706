"""GPU: ``bdrt_qp_bound`` (the replacement of cvxopt.solvers.qp inside Inverter._convex_opt, inversion.py:1043-1067) on
the 55 quadratic programs of the reference's own saved cvxopt runs (tests/golden/cvxopt_ridge.npz, see
test_oracle_cvxopt_ridge.py): one batch, every program solved by the CUDA block-principal-pivoting kernel.

  * programs it finishes at the tight tolerances (the same ones the oracle does): KKT residual below 1e-8 of |q|, objective
    equal to the oracle's exact one to 1e-10, never above cvxopt's and within cvxopt's own duality gap of it;
  * the near-singular rest (condition 1e11 .. 1e17): finished after the sign tolerances were relaxed, objective within
    2e-4 relative of cvxopt's."""
import numpy as np
import pytest
import torch

from oracle import ridge as oridge
from test_oracle_cvxopt_ridge import NAMES, programs

pytestmark = pytest.mark.gpu


def test_cuda_qp_on_the_references_cvxopt_programs():
    from bayes_drt_b200 import capi
    progs = [pr for name in NAMES for pr in programs(name)[1]]
    assert len(progs) == 55
    n = len(progs[0][1])
    P = torch.tensor(np.stack([pr[0] for pr in progs]))
    q = torch.tensor(np.stack([pr[1] for pr in progs]))
    x, kkt, iters = capi.qp_bound(P, q, torch.zeros(n, dtype=torch.float64))
    x, kkt, iters = x.cpu().numpy(), kkt.cpu().numpy(), iters.cpu().numpy()
    tight = 0
    for k, (Pk, qk, c, f_cvx, gap, _) in enumerate(progs):
        f = lambda v: 0.5 * v @ Pk @ v + qk @ v  # noqa: E731
        scale = np.max(np.abs(qk))
        assert x[k].min() >= 0.0
        xo, _, _, ito = oridge.qp_bound(Pk, qk, np.zeros(n))
        if iters[k] <= 100:
            tight += 1
            assert kkt[k] <= 1e-8 * scale, (k, kkt[k], scale)
            assert f(x[k]) <= f_cvx + 1e-12 * abs(f_cvx)
            assert f_cvx - f(x[k]) <= 1.05 * gap
            if ito <= 100:
                assert abs(f(x[k]) - f(xo)) <= 1e-10 * abs(f(xo))
        else:
            assert abs(f(x[k]) - f_cvx) <= 2e-4 * abs(f_cvx), (k, f(x[k]), f_cvx)
    assert tight >= 40, tight
