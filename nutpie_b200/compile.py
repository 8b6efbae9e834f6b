"""Front-end compilers of the reference, kept as named entry points.

`nutpie.compile_pymc_model` lowers a PyMC model to a numba cfunc
(python/nutpie/compile_pymc.py:523-624) and `nutpie.compile_stan_model` builds a
BridgeStan library (compile_stan.py:133-386); both yield HOST densities.  The
B200 engine evaluates densities on the device (nutpie_b200/csrc/models.cuh), and
neither pymc/pytensor nor bridgestan/stanc exist in this image, so the graph →
CUDA lowering is not part of this round; `from_cuda_source` is the hook it would target
(SURVEY.md §8f-3).  The functions
recognise the models that have a device density and otherwise explain what to
use instead — they never fall back to CPU sampling.
"""
from __future__ import annotations

from . import models


def compile_pymc_model(model=None, **kwargs):
    try:
        import pymc  # noqa: F401
    except ImportError as exc:
        raise ImportError(
            "pymc is not installed. The B200 engine samples device densities: build one with "
            "nutpie_b200.radon_model(...), normal_model(...) or funnel_model(...).") from exc
    raise NotImplementedError(
        "Lowering arbitrary PyMC graphs to CUDA is not implemented yet (SURVEY.md §8f-3); "
        "use a device density from nutpie_b200.models.")


def compile_stan_model(*, code=None, filename=None, **kwargs):
    raise NotImplementedError(
        "BridgeStan models are host densities and stanc is not available; the Stan example "
        "`x ~ normal(mu, 1)` of README.md:148-163 is nutpie_b200.normal_model(1, mu=mu).")


def from_pyfunc(*args, **kwargs):
    raise NotImplementedError(
        "Python-callable densities run on the host and the B200 engine never calls back into "
        "the host per gradient. Hand the density over as CUDA source instead: "
        "nutpie_b200.from_cuda_source(ndim, cuda_source, data=...).")


def from_cuda_source(ndim, cuda_source, data=None, *, scratch=0, shapes=None, dims=None,
                     coords=None):
    """The device counterpart of `from_pyfunc` (python/nutpie/compiled_pyfunc.py:108-155): the
    user supplies the log-density as CUDA C++ (`nb200_user_logp`, see include/nutpie_b200.h);
    it is compiled with NVRTC into the sampler kernel when the sampler is created."""
    return models.custom_model(ndim, cuda_source, data, scratch=scratch, shapes=shapes, dims=dims,
                               coords=coords)
