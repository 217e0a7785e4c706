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


@pytest.fixture(params=['toeplitz', 'toeplitz-generic', 'dense'])
def resident_A(request, monkeypatch):
    """Run a GPU test on every code path of the engine: the Toeplitz tables (picked automatically for shared log-uniform
    grids, two CTAs per SM) with the register-tiled per-slot phases, the same with the generic per-slot phases
    (BDRT_FORCE_GENERIC=1), and the dense-resident A (BDRT_FORCE_DENSE=1)."""
    monkeypatch.delenv('BDRT_FORCE_DENSE', raising=False)
    monkeypatch.delenv('BDRT_FORCE_GENERIC', raising=False)
    if request.param == 'dense':
        monkeypatch.setenv('BDRT_FORCE_DENSE', '1')
    elif request.param == 'toeplitz-generic':
        monkeypatch.setenv('BDRT_FORCE_GENERIC', '1')
    return request.param
