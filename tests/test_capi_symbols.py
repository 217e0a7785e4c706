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
    for cname, cls in (('bdrt_series_info', _lib.SeriesInfo), ('bdrt_series_data', _lib.SeriesData),
                       ('bdrt_lbfgs_opts', _lib.LbfgsOpts),
                       ('bdrt_newton_opts', _lib.NewtonOpts), ('bdrt_nuts_opts', _lib.NutsOpts),
                       ('bdrt_ridge_opts', _lib.RidgeOpts)):
        assert _struct_fields(hdr, cname) == [f[0] for f in cls._fields_], cname


def test_header_is_plain_c_and_links_from_c(lib, tmp_path):
    """The boundary is a C ABI: include/bdrt.h compiles as strict C99 and a C program links against libbdrt.so and calls
    the entry points that need no device (version, default options, sizes)."""
    import shutil
    import subprocess
    from bayes_drt_b200 import _lib
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    src = tmp_path / 'use_bdrt.c'
    src.write_text('''
#include <stdio.h>
#include <string.h>
#include "bdrt.h"
int main(void) {
  bdrt_lbfgs_opts lo; bdrt_nuts_opts no; bdrt_ridge_opts ro; bdrt_newton_opts wo; bdrt_series_data d;
  bdrt_lbfgs_default_opts(&lo); bdrt_nuts_default_opts(&no); bdrt_ridge_default_opts(&ro); bdrt_newton_default_opts(&wo);
  memset(&d, 0, sizeof d);
  d.model = BDRT_MODEL_SERIES | BDRT_MODEL_OUTLIERS; d.Nf = 81; d.K = 101; d.B = 1;
  printf("%d %d %d %d %g %d %g %d\\n", bdrt_version(), lo.history, no.chains, no.warmup, no.adapt_delta, ro.max_iter,
         ro.hl_beta, bdrt_num_params(&d));
  return 0;
}
''')
    exe = tmp_path / 'use_bdrt'
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-Werror', '-pedantic', '-I', os.path.join(ROOT, 'include'),
                    str(src), '-L', libdir, '-lbdrt', '-Wl,-rpath,' + libdir, '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    # Stan's defaults as the reference uses them (history 5; 2 chains, 200 warm-up, adapt_delta 0.9; ridge 20 / 2.5);
    # Series_outliers: D = 2 K + 9 + 2 Nf = 373
    assert out == ['100', '5', '2', '200', '0.9', '20', '2.5', '373']
