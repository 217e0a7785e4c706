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


@pytest.fixture(params=['toeplitz', 'toeplitz-generic', 'toeplitz-coop', 'dense'])
def resident_A(request, monkeypatch):
    """Run a GPU test on every code path of the engine: the Toeplitz tables with warp-private Hankel products (picked
    automatically for shared log-uniform grids, two CTAs per SM) and the register-tiled per-slot phases, the same with
    the generic per-slot phases (BDRT_FORCE_GENERIC=1), the Toeplitz tables with the cooperative CTA-wide products
    (BDRT_COOP=1; what the Newton kernel always uses), and the dense-resident A (BDRT_FORCE_DENSE=1)."""
    for k in ('BDRT_FORCE_DENSE', 'BDRT_FORCE_GENERIC', 'BDRT_COOP'):
        monkeypatch.delenv(k, raising=False)
    if request.param == 'dense':
        monkeypatch.setenv('BDRT_FORCE_DENSE', '1')
    elif request.param == 'toeplitz-generic':
        monkeypatch.setenv('BDRT_FORCE_GENERIC', '1')
    elif request.param == 'toeplitz-coop':
        monkeypatch.setenv('BDRT_COOP', '1')
    return request.param
