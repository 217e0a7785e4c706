"""CPU: the C-ABI library builds, loads, and exports every symbol include/bdrt.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from bayes_drt_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, 'include', 'bdrt.h')).read()
    declared = set(re.findall(r'\b(bdrt_[A-Za-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 20
    for s in declared:
        assert hasattr(lib, s), f'{s} declared in bdrt.h but not exported'
    from bayes_drt_b200 import _lib
    assert declared == set(_lib.SYMBOLS)


def test_version(lib):
    assert lib.bdrt_version() == 100


def test_no_cpu_fallback():
    import torch
    from bayes_drt_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_lib.BdrtError):
        _lib.Context()
