import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(ROOT, 'tests', 'golden')


@pytest.fixture(params=['toeplitz', 'toeplitz-generic', 'toeplitz-coop', 'dense', 'dense-global'])
def resident_A(request, monkeypatch):
    """Run a GPU test on every code path of the engine: the Toeplitz tables with warp-private Hankel products
    (BDRT_WARP=1 makes every kernel that has them use them; NUTS, per-spectrum grids and the log_prob hook do by
    default) and the register-tiled per-slot phases, the same with the generic per-slot phases (BDRT_FORCE_GENERIC=1),
    the Toeplitz tables with the cooperative CTA-wide products (BDRT_COOP=1; the default of the L-BFGS driver on a shared
    grid, and what the Newton kernel always uses), the dense-resident A (BDRT_FORCE_DENSE=1), and dense operands read from
    padded global copies (BDRT_FORCE_GDENSE=1: what two- / three-distribution models on general grids fall back to when
    their matrices do not fit in shared memory)."""
    for k in ('BDRT_FORCE_DENSE', 'BDRT_FORCE_GENERIC', 'BDRT_COOP', 'BDRT_WARP', 'BDRT_FORCE_GDENSE'):
        monkeypatch.delenv(k, raising=False)
    if request.param == 'dense':
        monkeypatch.setenv('BDRT_FORCE_DENSE', '1')
    elif request.param == 'dense-global':
        monkeypatch.setenv('BDRT_FORCE_DENSE', '1')
        monkeypatch.setenv('BDRT_FORCE_GDENSE', '1')
    elif request.param == 'toeplitz-generic':
        monkeypatch.setenv('BDRT_FORCE_GENERIC', '1')
        monkeypatch.setenv('BDRT_WARP', '1')
    elif request.param == 'toeplitz-coop':
        monkeypatch.setenv('BDRT_COOP', '1')
    else:
        monkeypatch.setenv('BDRT_WARP', '1')
    return request.param
