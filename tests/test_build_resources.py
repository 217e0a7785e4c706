"""CPU: register / spill budget of the solver kernels, read from the ptxas statistics of the in-tree build
(bayes_drt_b200/csrc/build.log, written by build.sh).  The two-CTAs-per-SM plan of the Toeplitz layouts
needs <= 128 registers per thread; the kernels sit exactly at that cap, and what they spill decides their speed
(profiles/r01i_ptxas_resources.txt).  A change that pushes the headline instantiations over these bounds should be noticed
before it reaches a GPU."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG = os.path.join(ROOT, 'bayes_drt_b200', 'csrc', 'build.log')


def _resources():
    rows, log = {}, open(LOG).read().split('\n')
    for i, l in enumerate(log):
        m = re.search(r"Function properties for (\S+)", l)
        if not m:
            continue
        sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", log[i + 1])
        reg = next((int(r.group(1)) for j in range(i + 1, min(i + 4, len(log)))
                    for r in [re.search(r"Used (\d+) registers", log[j])] if r), None)
        if sp and reg is not None:
            rows[m.group(1)] = (reg, int(sp.group(2)), int(sp.group(3)))
    names = subprocess.run(['c++filt'] + list(rows), capture_output=True, text=True).stdout.split('\n')
    return {re.sub(r'\(.*', '', n).replace('void ', ''): v for n, v in zip(names, rows.values())}


@pytest.mark.skipif(not os.path.exists(LOG), reason='no in-tree build log (run __graft_entry__.build())')
def test_solver_kernels_fit_two_ctas_per_sm():
    res = _resources()
    assert 'sm_100a' in open(LOG).read()
    # <TOEP, MK, FAST>: Toeplitz layouts (TOEP = 1 cooperative, 2 warp mode) run two CTAs of 256 threads per SM
    for kern in ('lbfgs_kernel', 'nuts_kernel', 'logpost_kernel'):
        inst = {k: v for k, v in res.items() if k.startswith(kern + '<') and not k.startswith(kern + '<0')}
        assert inst, kern
        for k, (reg, st, ld) in inst.items():
            assert reg <= 128, (k, reg)
    # the benchmarked instantiations: spills stay where they were measured (bytes per thread)
    for k, max_ld in (('lbfgs_kernel<1, 0, 1>', 900), ('nuts_kernel<2, 0, 1>', 1700), ('logpost_kernel<2, 0, 1>', 600)):
        assert k in res, k
        assert res[k][2] <= max_ld, (k, res[k])
    # dense-resident instantiations take the whole SM (one CTA): no spills at 255 registers
    for k, (reg, st, ld) in res.items():
        if k.startswith(('lbfgs_kernel<0', 'logpost_kernel<0')):
            assert st == 0 and ld == 0, (k, reg, st, ld)
