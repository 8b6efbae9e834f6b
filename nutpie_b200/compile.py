"""Front-end entry points: how a density gets to the B200 engine.

The reference has three producers of a HOST density —
`compile_pymc_model` (PyTensor graph -> numba cfunc, python/nutpie/compile_pymc.py:523-624),
`compile_stan_model` (BridgeStan shared object, compile_stan.py:133-386) and
`from_pyfunc` (Python callables, compiled_pyfunc.py:108-155) — and hands nuts-rs the function
pointer (src/pymc.rs:50-62) or the callable (src/pyfunc.rs:34-84).  The engine here takes

  * device densities (nutpie_b200.models: normal / funnel / radon, hand-written CUDA),
  * `from_cuda_source`: the density as CUDA C++, compiled with NVRTC into the sampler kernel,
  * the reference's own plug-in ABI, unchanged: `from_cfuncs` (any C pointer with the
    RawLogpFunc / RawExpandFunc signatures, e.g. numba cfuncs — exactly what
    `compile_pymc_model` produces) and `from_pyfunc` (Python callables).  These sample through
    the NB200_MODEL_HOST kind: the persistent kernel posts each position to a mailbox in mapped
    pinned memory and host threads call the pointer (include/nutpie_b200.h).  Slow — one host
    call per gradient, like the reference — but every existing LogpFunc runs.

`compile_pymc_model` itself needs pymc + pytensor to turn a model graph into that cfunc; neither
exists in this image, so it is implemented only up to the point that can run here (see its
docstring).  BridgeStan is out of scope.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from functools import partial
from typing import Any, Callable

import numpy as np

from . import _lib, models


@dataclass(frozen=True)
class CompiledModel:
    """python/nutpie/sample.py:17-60 — what `sample()` needs from any compiled model."""

    dims: dict | None
    reparameterized_names: list | None = field(default=None, kw_only=True)

    @property
    def n_dim(self) -> int:
        raise NotImplementedError()

    @property
    def shapes(self):
        raise NotImplementedError()

    @property
    def coords(self):
        raise NotImplementedError()

    def _make_sampler(self, *args, **kwargs):
        raise NotImplementedError()

    def _make_model(self, *args, **kwargs):
        raise NotImplementedError()

    # host models: the trace holds unconstrained draws, expanded through the model object
    def _expand(self, q):
        return self._model_for_expand()._expand(q)

    def _split_expanded(self, e):
        return self._model_for_expand()._split_expanded(e)

    def _model_for_expand(self):
        m = self.__dict__.get("_cached_model")
        if m is None:
            m = self._make_model(None)
            object.__setattr__(self, "_cached_model", m)
        return m


@dataclass(frozen=True)
class CompiledCFuncModel(CompiledModel):
    """The shape of CompiledPyMCModel (python/nutpie/compile_pymc.py:104-236) without the PyMC
    graph: compiled logp / expand C functions with `.address` (numba `CFunc`s or ctypes function
    objects), the `user_data` record they read, the expanded-variable table."""

    compiled_logp_func: Any
    compiled_expand_func: Any
    initial_point_func: Callable[[int], np.ndarray]
    user_data: np.ndarray | None
    n_expanded: int
    shape_info: Any  # (names, slices-or-None, shapes) as in compile_pymc.py:206-209
    _n_dim: int
    _shapes: dict
    _coords: dict | None

    @property
    def n_dim(self):
        return self._n_dim

    @property
    def shapes(self):
        return self._shapes

    @property
    def coords(self):
        return self._coords or {}

    def with_user_data(self, user_data):
        """compile_pymc.py:136-161 (`with_data`): the same compiled functions on new data."""
        return dataclasses.replace(self, user_data=user_data)

    def _make_sampler(self, settings, init_mean, cores, progress_type, extra_callback=None,
                      extra_callback_rate=500, store=None, **kw):
        model = self._make_model(init_mean)  # compile_pymc.py:168-187
        return _lib.PySampler.from_pymc(settings, cores, model, progress_type, extra_callback,
                                        extra_callback_rate, store, **kw)

    def _make_model(self, init_mean):
        """compile_pymc.py:189-233, call for call."""
        ud = 0 if self.user_data is None else self.user_data.ctypes.data
        expand_fn = _lib.ExpandFunc(self.n_dim, self.n_expanded, _address(self.compiled_expand_func),
                                    ud, self)
        logp_fn = _lib.LogpFunc(_address(self.compiled_logp_func), ud, self)
        var_names = self.shape_info[0]
        coords = dict(self._coords) if self._coords is not None else {}
        dim_sizes = {name: len(vals) for name, vals in coords.items()}
        dims = dict(self.dims) if self.dims is not None else {}
        var_types = ["float64"] * len(var_names)
        var_shapes = self.shape_info[2]
        variables = _lib.PyVariable.new_variables(var_names, var_types, var_shapes, dim_sizes, dims)
        return _lib.PyMcModel(logp_fn, expand_fn, variables, self.n_dim, dim_sizes, coords,
                              self.initial_point_func, None)


def _address(f) -> int:
    if hasattr(f, "address"):  # numba CFunc
        return int(f.address)
    import ctypes as C

    return int(C.cast(f, C.c_void_p).value)


def from_cfuncs(n_dim: int, logp_cfunc, expand_cfunc, expanded_names, expanded_shapes, *,
                initial_point_fn=None, user_data=None, dims=None, coords=None,
                reparameterized_names=None) -> CompiledCFuncModel:
    """A compiled model from C function pointers in the reference's plug-in ABI — what
    `compile_pymc_model` ends with (`_make_c_logp_func` / `_make_c_expand_func`,
    compile_pymc.py:970-1043):

        int logp(size_t dim, const double *x, double *grad, double *logp, const void *user_data)
        int expand(size_t dim, size_t n_expanded, const double *x, double *out, const void *ud)

    `initial_point_fn(seed) -> float64[n_dim]` as PyMC's `make_initial_point_fn`
    (compile_pymc.py:596-604); default: zeros + U(-1, 1) jitter like PyMC's default init."""
    n_dim = int(n_dim)
    shapes = [tuple(int(x) for x in s) for s in expanded_shapes]
    n_expanded = int(sum(int(np.prod(s)) if s else 1 for s in shapes))
    if initial_point_fn is None:
        def initial_point_fn(seed):
            return np.random.default_rng(seed).uniform(-1.0, 1.0, n_dim)
    return CompiledCFuncModel(
        dims=dict(dims or {}), reparameterized_names=list(reparameterized_names or []),
        compiled_logp_func=logp_cfunc, compiled_expand_func=expand_cfunc,
        initial_point_func=initial_point_fn, user_data=user_data, n_expanded=n_expanded,
        shape_info=(list(expanded_names), None, shapes), _n_dim=n_dim,
        _shapes=dict(zip(expanded_names, shapes)), _coords=dict(coords or {}))


@dataclass(frozen=True)
class PyFuncModel(CompiledModel):
    """python/nutpie/compiled_pyfunc.py:14-105."""

    _make_logp_func: Callable
    _make_expand_func: Callable
    _make_initial_points: Callable[[int], np.ndarray] | None
    _shared_data: dict
    _n_dim: int
    _variables: list
    _dim_sizes: dict
    _coords: dict
    _raw_logp_fn: Callable | None = None

    @property
    def shapes(self):
        return {var.name: tuple(var.shape) for var in self._variables}

    @property
    def coords(self):
        return self._coords

    @property
    def n_dim(self):
        return self._n_dim

    def with_data(self, **updates):
        for name in updates:
            if name not in self._shared_data:
                raise ValueError(f"Unknown data variable: {name}")
        updated = self._shared_data.copy()
        updated.update(**updates)
        return dataclasses.replace(self, _shared_data=updated)

    def with_transform_adapt(self, **kwargs):
        raise NotImplementedError("adaptation='flow' is not supported by the B200 engine")

    def _make_sampler(self, settings, init_mean, cores, progress_type, extra_callback=None,
                      extra_callback_rate=500, store=None, **kw):
        model = self._make_model(init_mean)
        return _lib.PySampler.from_pyfunc(settings, cores, model, progress_type, extra_callback,
                                          extra_callback_rate, store, **kw)

    def _make_model(self, init_mean):
        def make_logp_func():
            logp_fn = self._make_logp_func()
            return partial(logp_fn, **self._shared_data)

        def make_expand_func(seed1, seed2, chain):
            expand_fn = self._make_expand_func(seed1, seed2, chain)
            return partial(expand_fn, **self._shared_data)

        return _lib.PyModel(make_logp_func, make_expand_func, self._variables, self.n_dim,
                            dim_sizes=self._dim_sizes, coords=self._coords,
                            init_point_func=self._make_initial_points, transform_adapter=None)


def from_pyfunc(ndim: int, make_logp_fn: Callable, make_expand_fn: Callable,
                expanded_dtypes: list, expanded_shapes: list, expanded_names: list, *,
                coords: dict | None = None, dims: dict | None = None,
                shared_data: dict | None = None, make_initial_point_fn=None,
                make_transform_adapter=None, raw_logp_fn=None, reparameterized_names=None):
    """python/nutpie/compiled_pyfunc.py:108-155, same arguments.  `make_logp_fn()` returns
    `logp(x, **shared_data) -> (logp, grad)`; `make_expand_fn(seed1, seed2, chain)` returns
    `expand(x, **shared_data) -> {name: array}`.  Sampled through the host plug-in service."""
    if make_transform_adapter is not None:
        raise NotImplementedError("adaptation='flow' is not supported by the B200 engine")
    coords = dict(coords or {})
    dims = dict(dims or {})
    shared_data = dict(shared_data or {})
    dim_sizes = {k: len(v) for k, v in coords.items()}
    shapes = [tuple(shape) for shape in expanded_shapes]
    variables = _lib.PyVariable.new_variables(
        expanded_names, [str(np.dtype(dtype)) for dtype in expanded_dtypes], shapes, dim_sizes, dims)
    return PyFuncModel(
        dims=dims, reparameterized_names=reparameterized_names, _n_dim=int(ndim), _coords=coords,
        _dim_sizes=dim_sizes, _make_logp_func=make_logp_fn, _make_expand_func=make_expand_fn,
        _make_initial_points=make_initial_point_fn, _variables=variables,
        _shared_data=shared_data, _raw_logp_fn=raw_logp_fn)


def compile_pymc_model(model=None, *, backend="numba", **kwargs):
    """python/nutpie/compile_pymc.py:523-624.  The reference lowers the model's logp graph with
    PyTensor to a numba cfunc and wraps it as `LogpFunc`; the engine accepts exactly that object
    (`from_cfuncs`, NB200_MODEL_HOST), so with pymc installed the reference's own
    `nutpie.compile_pymc._compile_pymc_model_numba` output can be passed to `from_cfuncs`
    unchanged.  The graph lowering itself is not re-implemented here: pymc / pytensor are not
    installable in this image, so it could not be run even once."""
    try:
        import pymc  # noqa: F401
        import pytensor  # noqa: F401
    except ImportError as exc:
        raise ImportError(
            "pymc/pytensor are not installed, so a PyMC graph cannot be lowered here. Hand the "
            "engine the compiled functions instead: nutpie_b200.from_cfuncs(n_dim, logp_cfunc, "
            "expand_cfunc, names, shapes) takes numba cfuncs with the reference's signatures "
            "(compile_pymc.py:975-981, 1018-1024); device densities are in "
            "nutpie_b200.models.") from exc
    raise NotImplementedError(
        "PyTensor graph lowering is not re-implemented: compile with the reference's "
        "compile_pymc_model and pass compiled_logp_func / compiled_expand_func / user_data / "
        "shape_info to nutpie_b200.from_cfuncs (same plug-in ABI, src/pymc.rs:23-37).")


def compile_stan_model(*, code=None, filename=None, **kwargs):
    raise NotImplementedError(
        "BridgeStan models need stanc + bridgestan, which are out of scope (SURVEY.md §2 N12); "
        "the Stan example `x ~ normal(mu, 1)` of README.md:148-163 is "
        "nutpie_b200.normal_model(1, mu=mu).")


def from_cuda_source(ndim, cuda_source, data=None, *, scratch=0, shapes=None, dims=None,
                     coords=None):
    """The device counterpart of `from_pyfunc` (python/nutpie/compiled_pyfunc.py:108-155): the
    user supplies the log-density as CUDA C++ (`nb200_user_logp`, see include/nutpie_b200.h);
    it is compiled with NVRTC into the sampler kernel when the sampler is created."""
    return models.custom_model(ndim, cuda_source, data, scratch=scratch, shapes=shapes, dims=dims,
                               coords=coords)
