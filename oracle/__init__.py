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
  (pystan==2.19.1.1, cvxopt).  The reference ships no test or golden vector
  for them, so for those parts the oracle is a restatement of the published
  algorithm anchored on the reference's call sites and on the loose paper
  outputs in ``code_EchemActa`` -- **parity unpinned** in the strict sense
  (see DESIGN.md).
"""
