"""CPU oracle for the bayes-drt inversion hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as
the thing timed *as the CPU baseline*.  The product path
(``bayes_drt_b200``) never imports this package and fails loudly when its CUDA
library is missing.

Every function cites the reference file:line it restates (paths relative to
the upstream repository root, ``/root/reference`` in the build container).

Parity pinning
--------------
* ``oracle.matrices`` is pinned against the reference's own ``matrices.py``
  (imported verbatim in the build container by
  ``scripts/make_golden_matrices.py``; outputs committed in ``tests/golden/matrices.npz``).
* The Stan programs, Stan's L-BFGS / NUTS and cvxopt's QP are third-party
  natives that are absent from the reference tree and from this image
  (pystan==2.19.1.1, cvxopt), and the reference ships no test for them.
  What the tree does hold are results its authors saved for their paper:
  - ``code_EchemActa/map_results/obj_*.pkl``: everything ``StanModel.optimizing``
    returned for 25 MAP fits -> the model arithmetic, the constants handed to
    Stan and the location of the optimum are pinned to Stan's own numbers
    (``tests/golden/stan_map.npz``, ``tests/test_oracle_stan_map.py``);
  - ``code_EchemActa/comparisons/hyper-ridge/results/obj_*.pkl``: every QP
    solution of nine hyper-lambda ridge fits with cvxopt's objective and gap ->
    the exact QP solver and the lambda update are pinned to cvxopt in the
    objective (``tests/golden/cvxopt_ridge.npz``,
    ``tests/test_oracle_cvxopt_ridge.py``).
  What stays **unpinned** is the path of Stan's optimiser and its sampler
  (restatements of the published algorithms, compared with the CUDA drivers
  statistically / iterate by iterate).
"""
