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


def _struct_fields(hdr, name):
    """field names of `typedef struct { ... } name;` in declaration order"""
    body = re.search(r'typedef struct \{([^{}]*)\}\s*' + name + r'\s*;', hdr, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    out = []
    for stmt in body.split(';'):
        stmt = stmt.strip()
        if not stmt:
            continue
        # "const double* A" | "double sigma_min, ups_alpha" | "double reg_ord[3]" | "int Nf"
        names = stmt.split(None, 1)[1] if not stmt.startswith(('const', 'unsigned', 'long long')) else \
            re.sub(r'^(const\s+\w+(\s+\w+)?\s*\*?|unsigned long long|long long)\s*', '', stmt)
        for nm in names.split(','):
            out.append(re.sub(r'[\*\s]|\[.*\]', '', nm))
    return out


def test_ctypes_structs_mirror_the_header():
    """The ctypes Structures the Python host passes by reference must have the header's fields in the header's order."""
    from bayes_drt_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'bdrt.h')).read()
    for cname, cls in (('bdrt_series_data', _lib.SeriesData), ('bdrt_lbfgs_opts', _lib.LbfgsOpts),
                       ('bdrt_newton_opts', _lib.NewtonOpts), ('bdrt_nuts_opts', _lib.NutsOpts),
                       ('bdrt_ridge_opts', _lib.RidgeOpts)):
        assert _struct_fields(hdr, cname) == [f[0] for f in cls._fields_], cname
