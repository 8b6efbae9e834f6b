"""nutpie_b200 — nutpie's sampling surface on a B200-native NUTS engine.

Public names follow python/nutpie/__init__.py:1-18 of the reference.  The
sampler core (what nutpie gets from nuts-rs) runs as sm_100a CUDA kernels behind
the C-ABI in include/nutpie_b200.h; models are device densities
(nutpie_b200.models) instead of host function pointers.
"""
from . import models
from ._lib import PyChainProgress as ChainProgress
from ._lib import __version__
from .compile import (compile_pymc_model, compile_stan_model, from_cfuncs, from_cuda_source,
                      from_pyfunc)
from .datasets import make_radon_data
from .models import custom_model, funnel_model, normal_model, radon_model
from .sample import Trace, sample

__all__ = [
    "__version__", "sample", "compile_pymc_model", "compile_stan_model", "from_pyfunc",
    "ChainProgress", "Trace", "models", "normal_model", "funnel_model", "radon_model",
    "make_radon_data", "from_cuda_source", "custom_model", "from_cfuncs",
]
