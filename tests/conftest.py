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


@pytest.fixture(params=['toeplitz', 'dense'])
def resident_A(request, monkeypatch):
    """Run a GPU test on both resident-operand layouts of the engine: the Toeplitz tables (picked automatically for
    shared log-uniform grids, two CTAs per SM) and the dense A (forced with BDRT_FORCE_DENSE=1)."""
    if request.param == 'dense':
        monkeypatch.setenv('BDRT_FORCE_DENSE', '1')
    else:
        monkeypatch.delenv('BDRT_FORCE_DENSE', raising=False)
    return request.param
