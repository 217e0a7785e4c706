"""ctypes binding of libbdrt.so (the C ABI declared in include/bdrt.h).

There is NO CPU fallback: if the CUDA library is missing, or no GPU is visible when a compute call is made, this
module raises.  PyTorch is only the container for device memory and streams.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('BDRT_LIB') or os.path.join(_HERE, 'libbdrt.so')  # BDRT_LIB: A/B timing of library builds

# symbols declared in include/bdrt.h (tests check that every one is exported)
SYMBOLS = [
    'bdrt_version', 'bdrt_ctx_create', 'bdrt_ctx_destroy', 'bdrt_last_error', 'bdrt_launch_count',
    'bdrt_build_A', 'bdrt_build_L', 'bdrt_build_M', 'bdrt_num_params', 'bdrt_num_outputs', 'bdrt_series_analyze',
    'bdrt_logpost_grad',
    'bdrt_lbfgs_default_opts', 'bdrt_map_lbfgs', 'bdrt_newton_default_opts', 'bdrt_map_newton',
    'bdrt_nuts_default_opts', 'bdrt_nuts', 'bdrt_constrain', 'bdrt_summarize', 'bdrt_diagnostics', 'bdrt_qp_bound', 'bdrt_ridge_default_opts',
    'bdrt_ridge_fit', 'bdrt_peak_fp64',
]

KERNEL = {'DRT': 0, 'DDT': 1}
DIST = {'series': 0, 'parallel': 1}
SYM = {'planar': 0, 'spherical': 1}
BC = {'transmissive': 0, 'blocking': 1}
MODEL_SERIES, MODEL_SERIES_PARALLEL, MODEL_PARALLEL, MODEL_SERIES_2PARALLEL, MODEL_POS, MODEL_OUTLIERS = 0, 1, 2, 3, 16, 32

TERM_NAMES = {0: 'running', 10: 'absx', 20: 'absf', 21: 'relf', 30: 'absgrad', 31: 'relgrad', 40: 'maxit',
              -1: 'lsfail', -2: 'badinit'}


class SeriesInfo(C.Structure):
    _fields_ = [('valid', C.c_int), ('toepA', C.c_int), ('bw', C.c_int * 3), ('toepL', C.c_int * 3),
                ('taps', C.c_double * 13 * 3 * 3)]


class SeriesData(C.Structure):
    _fields_ = [('model', C.c_int), ('Nf', C.c_int), ('K', C.c_int), ('B', C.c_int), ('per_spectrum_grid', C.c_int),
                ('A', C.c_void_p), ('Z', C.c_void_p), ('freq', C.c_void_p), ('L', C.c_void_p),
                ('sigma_min', C.c_double), ('ups_alpha', C.c_double), ('ups_beta', C.c_double),
                ('induc_scale', C.c_double), ('sigma_out_lambda', C.c_double), ('sigma_out_alpha', C.c_double),
                ('sigma_out_beta', C.c_double),
                ('Kp', C.c_int), ('Ap', C.c_void_p), ('Lp', C.c_void_p), ('x_sum_invscale', C.c_double),
                ('xp_scale', C.c_double), ('Kp2', C.c_int), ('Ap2', C.c_void_p), ('Lp2', C.c_void_p),
                ('xp2_scale', C.c_double), ('info', C.POINTER(SeriesInfo))]


class LbfgsOpts(C.Structure):
    _fields_ = [('max_iter', C.c_int), ('history', C.c_int), ('init_alpha', C.c_double), ('tol_obj', C.c_double),
                ('tol_rel_obj', C.c_double), ('tol_grad', C.c_double), ('tol_rel_grad', C.c_double),
                ('tol_param', C.c_double)]


class NewtonOpts(C.Structure):
    _fields_ = [('max_iter', C.c_int), ('gtol', C.c_double), ('fd_step', C.c_double)]


class NutsOpts(C.Structure):
    _fields_ = [('chains', C.c_int), ('warmup', C.c_int), ('samples', C.c_int), ('max_treedepth', C.c_int),
                ('adapt_delta', C.c_double), ('adapt_t0', C.c_double), ('adapt_gamma', C.c_double),
                ('adapt_kappa', C.c_double), ('seed', C.c_ulonglong), ('spectrum_offset', C.c_longlong),
                ('spectrum_ids', C.c_void_p)]


class RidgeOpts(C.Structure):
    _fields_ = [('penalty', C.c_int), ('nonneg', C.c_int), ('max_iter', C.c_int), ('xtol', C.c_double),
                ('hl_beta', C.c_double), ('lambda_0', C.c_double), ('reg_ord', C.c_double * 3),
                ('L1_penalty', C.c_double), ('epsilon', C.c_double), ('fit_inductance', C.c_int),
                ('hl_fbeta', C.c_double), ('stop_rule', C.c_int)]


class BdrtError(RuntimeError):
    pass


_lib = None


def load():
    """Load libbdrt.so (built in-tree by __graft_entry__.build() / csrc/build.sh).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BdrtError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                        f'(bayes_drt_b200 has no CPU fallback)')
    lib = C.CDLL(LIB_PATH)
    lib.bdrt_last_error.restype = C.c_char_p
    lib.bdrt_launch_count.restype = C.c_longlong
    _lib = lib
    return lib


class Context:
    """One bdrt_ctx bound to a CUDA device and to torch's current stream on that device."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise BdrtError('bayes_drt_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        lib = load()
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.lib = lib
        self._h = C.c_void_p()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = lib.bdrt_ctx_create(self.device.index, C.c_void_p(stream), C.byref(self._h))
        if rc != 0:
            raise BdrtError(f'bdrt_ctx_create failed with status {rc}')

    def __del__(self):
        try:
            if getattr(self, '_h', None) and self._h.value:
                self.lib.bdrt_ctx_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            msg = self.lib.bdrt_last_error(self._h)
            raise BdrtError(f'libbdrt status {rc}: {msg.decode() if msg else ""}')

    @property
    def launches(self):
        return int(self.lib.bdrt_launch_count(self._h))


_contexts = {}


def context(device=None):
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device())
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    if key not in _contexts:
        _contexts[key] = Context(dev)
    return _contexts[key]


def ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def f64(t, device):
    """contiguous float64 tensor on device (no copy when already so)"""
    return torch.as_tensor(t, dtype=torch.float64, device=device).contiguous()
