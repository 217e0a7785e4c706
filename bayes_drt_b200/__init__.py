"""bayes_drt_b200: B200-native (sm_100a) implementation of the bayes-drt inversion hot path.

Python host code over a C-ABI CUDA library (libbdrt.so, include/bdrt.h); PyTorch tensors are the batch container.
There is no CPU fallback: compute calls raise if the library or the GPU is missing.
"""
from . import _lib  # noqa: F401

__all__ = ['Inverter', 'matrices', 'capi', 'synth']


def __getattr__(name):
    if name == 'Inverter':
        from .inverter import Inverter
        return Inverter
    if name in ('matrices', 'capi', 'synth', 'distributed'):
        import importlib
        return importlib.import_module('.' + name, __name__)
    raise AttributeError(name)
