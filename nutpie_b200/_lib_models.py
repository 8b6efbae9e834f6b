"""The model-side classes of `nutpie._lib` that the reference's Python layer constructs.

  PyVariable.new_variables   src/common.rs:283-300, 381-465
  LogpFunc / ExpandFunc      src/pymc.rs:39-98     (raw C pointers + user_data + keep-alive)
  PyMcModel                  src/pymc.rs:410-472   (what CompiledPyMCModel._make_model builds,
                                                    python/nutpie/compile_pymc.py:189-233)
  PyModel                    src/pyfunc.rs:21-84   (what PyFuncModel._make_model builds,
                                                    python/nutpie/compiled_pyfunc.py:72-105)
  store.*                    src/wrapper.rs:1754-1756 (referenced at import time by
                                                    python/nutpie/sample.py:472-478)

In the reference these objects end up inside `nuts_rs::Sampler`, whose worker threads call the
density pointer once per leapfrog (src/pymc.rs:197-215).  Here they describe a
NB200_MODEL_HOST model (include/nutpie_b200.h): the sampler still runs as the persistent CUDA
kernel, and the pointer is called by the engine's host service threads through a mailbox in
mapped pinned memory.  Imported into the `_lib` namespace at the bottom of _lib.py.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

_ITEM_TYPES = {"uint64": np.uint64, "int64": np.int64, "float64": np.float64,
               "float32": np.float32, "bool": np.bool_, "string": object}


class PyVariable:
    """src/common.rs:283-300 — one named slice of the expanded vector."""

    __slots__ = ("name", "item_type", "dims", "shape", "num_elements", "start_idx", "end_idx")

    def __init__(self, name, item_type, shape, all_dims, dim_sizes, start_idx):
        # PyVariable::new, src/common.rs:302-379: dims and shape complete each other
        dims = all_dims.get(name)
        if dims is not None and shape is not None:
            if len(dims) != len(shape):
                raise RuntimeError(
                    f"Variable '{name}': number of dims ({len(dims)}) does not match number of "
                    f"shape entries ({len(shape)})")
            for dim, size in zip(dims, shape):
                if dim in dim_sizes and dim_sizes[dim] != size:
                    raise RuntimeError(
                        f"Variable '{name}': dimension '{dim}' has inconsistent size. Expected "
                        f"{size}, but previously defined as {dim_sizes[dim]}")
            dims = list(dims)
        elif dims is not None:
            shape = []
            for dim in dims:
                if dim not in dim_sizes:
                    raise RuntimeError(f"Variable '{name}': dimension '{dim}' size unknown and "
                                       "no shape provided")
                shape.append(dim_sizes[dim])
            dims = list(dims)
        elif shape is not None:
            dims = []
            for i, size in enumerate(shape):
                gen = f"{name}_dim_{i}"
                if gen in dim_sizes:
                    raise RuntimeError(f"Variable '{name}': generated anonymous dimension name "
                                       f"'{gen}' already exists.")
                dim_sizes[gen] = int(size)
                dims.append(gen)
            all_dims[name] = list(dims)
        else:
            raise RuntimeError(f"Variable '{name}': no dims or shape provided")
        self.name = name
        self.item_type = item_type
        self.dims = list(dims)
        self.shape = [int(x) for x in shape]
        self.num_elements = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        self.start_idx = start_idx
        self.end_idx = None if start_idx is None else start_idx + self.num_elements

    @property
    def dtype(self):
        return np.dtype(_ITEM_TYPES[self.item_type])

    @classmethod
    def new_variables(cls, names, item_types, shapes, dim_sizes, dims):
        """src/common.rs:384-465: consecutive slices of the expanded vector; `dim_sizes` and
        `dims` (the caller's dicts) receive the anonymous dimensions that were generated."""
        all_dims = {str(k): [str(x) for x in v] for k, v in dims.items()}
        sizes = {str(k): int(v) for k, v in dim_sizes.items()}
        out, pos = [], 0
        for name, item_type, shape in zip(names, item_types, shapes):
            if item_type not in _ITEM_TYPES:
                raise RuntimeError(f"Unsupported item type: {item_type}")
            shape = None if shape is None else [int(x) for x in shape]
            try:
                var = cls(str(name), item_type, shape, all_dims, sizes, pos)
            except RuntimeError as exc:
                raise RuntimeError(f"Could not create variable: {exc}") from None
            pos += var.num_elements
            out.append(var)
        for k, v in sizes.items():
            if k not in dim_sizes:
                dim_sizes[k] = v
        for k, v in all_dims.items():
            if k not in dims:
                dims[k] = list(v)
        return out

    def __repr__(self):
        return (f"PyVariable(name={self.name!r}, dims={self.dims}, shape={self.shape}, "
                f"start_idx={self.start_idx}, end_idx={self.end_idx})")


class LogpFunc:
    """src/pymc.rs:39-62 — `ptr` is the address of a C function
    `int logp(size_t dim, const double* x, double* grad, double* logp, const void* user_data)`
    (e.g. a numba cfunc's `.address`, compile_pymc.py:197-201)."""

    def __init__(self, ptr, user_data_ptr, keep_alive):
        self.ptr, self.user_data_ptr, self._keep_alive = int(ptr), int(user_data_ptr or 0), keep_alive


class ExpandFunc:
    """src/pymc.rs:64-95 — `int expand(size_t dim, size_t expanded_dim, const double* x,
    double* out, const void* user_data)`."""

    def __init__(self, dim, expanded_dim, ptr, user_data_ptr, keep_alive):
        self.dim, self.expanded_dim = int(dim), int(expanded_dim)
        self.ptr, self.user_data_ptr, self._keep_alive = int(ptr), int(user_data_ptr or 0), keep_alive


def _check_dicts(dim_sizes, coords):
    for k, v in dim_sizes.items():
        if not isinstance(k, str):
            raise RuntimeError("Dimension key is not a string")
        if not isinstance(v, (int, np.integer)):
            raise RuntimeError("Dimension size value is not an integer")
    for k in coords:
        if not isinstance(k, str):
            raise RuntimeError("Coordinate key is not a string")


def chain_seed(seed: int, chain: int) -> int:
    """The u64 handed to `init_func(seed)` for one chain.  nuts-rs takes `rng.next_u64()` from
    the chain's ChaCha stream (src/pymc.rs:510); this engine's streams are Philox counters, so
    the seed is a splitmix64 hash of (settings.seed, global chain id) — deterministic, distinct
    per chain, independent of how chains are sharded over GPUs."""
    z = (int(seed) + 0x9E3779B97F4A7C15 * (int(chain) + 1)) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


class _HostModelBase:
    """What PySampler needs from a host model: a NB200_MODEL_HOST descriptor, the variable table
    and `expand`."""

    kind = "host"
    reparameterized_names: list = []

    @property
    def n_dim(self):
        return self.dim

    def _variable_dims(self):
        return {v.name: tuple(v.dims) for v in self.variables}

    def _variable_shapes(self):
        return {v.name: tuple(v.shape) for v in self.variables}

    def _initial_points(self, seed, n_chains, chain_id_offset):
        f = self.init_func
        if f is None:
            return None
        q0 = np.empty((n_chains, self.dim))
        for c in range(n_chains):
            pt = np.asarray(f(chain_seed(seed, chain_id_offset + c)))
            if pt.dtype != np.float64 or pt.ndim != 1:
                raise RuntimeError("Initialization array returned incorrect argument")
            if not pt.flags["C_CONTIGUOUS"]:
                raise RuntimeError("Initial point must be contiguous")
            if pt.shape[0] != self.dim:
                raise RuntimeError("Initial point has incorrect length")
            q0[c] = pt
        return q0

    def _split_expanded(self, e):
        out = {}
        for v in self.variables:
            a = e[..., v.start_idx:v.end_idx]
            a = a.reshape(e.shape[:-1] + tuple(v.shape)) if v.shape else a[..., 0]
            if v.item_type == "bool":
                a = a != 0.0
            elif v.item_type not in ("float64", "string"):
                a = a.astype(_ITEM_TYPES[v.item_type])
            out[v.name] = a
        return out


class PyMcModel(_HostModelBase):
    """src/pymc.rs:410-472 — density + expand pointers, variable table, `init_func(seed)`."""

    def __init__(self, density, expand, variables, dim, dim_sizes, coords, init_func,
                 transform_adapter=None):
        if not isinstance(density, LogpFunc) or not isinstance(expand, ExpandFunc):
            raise TypeError("PyMcModel needs a LogpFunc and an ExpandFunc")
        _check_dicts(dim_sizes, coords)
        self.density, self.expand = density, expand
        self.variables = list(variables)
        self.dim = int(dim)
        self.dim_sizes = dict(dim_sizes)
        self.coords = dict(coords)
        self.init_func = init_func
        self.transform_adapter = transform_adapter  # flow adaptation: out of scope
        self.host_threads = 0

    def _descriptor(self):
        from . import _lib

        d = _lib.ModelDesc()
        d.kind = _lib.MODEL_KINDS["host"]
        d.dim = self.dim
        d.sigma = 1.0
        d.host_logp = self.density.ptr
        d.host_user_data = self.density.user_data_ptr or None
        d.host_expand = self.expand.ptr
        d.host_expand_user_data = self.expand.user_data_ptr or None
        d.host_expanded_dim = self.expand.expanded_dim
        d.host_threads = int(self.host_threads)
        return d, [self.density, self.expand]

    def _expand(self, q):
        """CpuLogpFunc::expand_vector (src/pymc.rs:217-286) for every stored draw: the expand
        pointer is called on host threads by the C library, then sliced by variable."""
        from . import _lib

        L = _lib.load_library()
        q = np.asarray(q, dtype=np.float64)
        lead = q.shape[:-1]
        if self.expand.dim != self.dim:
            raise RuntimeError("Expand function returned error code -1")
        n = int(np.prod(lead, dtype=np.int64)) if lead else 1
        out = np.empty((n, self.expand.expanded_dim))
        flat = np.ascontiguousarray(q).reshape(n, self.dim)
        if n:
            _lib._check(L.nb200_host_expand_rows(
                C.c_void_p(self.expand.ptr), C.c_void_p(self.expand.user_data_ptr or None),
                self.dim, self.expand.expanded_dim, n, C.c_void_p(flat.ctypes.data), self.dim,
                C.c_void_p(out.ctypes.data), 0))
        return self._split_expanded(out.reshape(lead + (self.expand.expanded_dim,)))


class PyModel(_HostModelBase):
    """src/pyfunc.rs:21-84 — the density is a Python callable `logp(x) -> (float, grad)` made by
    `make_logp_func()`; `make_expand_func(seed1, seed2, chain)` makes `expand(x) -> dict`.
    The callable is wrapped in a C trampoline with the plug-in signature; it runs under the
    GIL on ONE host service thread (src/pyfunc.rs:206-230 holds the GIL per call as well)."""

    def __init__(self, make_logp_func, make_expand_func, variables, ndim, dim_sizes, coords, *,
                 init_point_func=None, transform_adapter=None):
        _check_dicts(dim_sizes, coords)
        self.make_logp_func, self.make_expand_func = make_logp_func, make_expand_func
        self.variables = list(variables)
        self.dim = int(ndim)
        self.dim_sizes, self.coords = dict(dim_sizes), dict(coords)
        self.init_func = init_point_func
        self.transform_adapter = transform_adapter
        self.host_threads = 1
        self.last_error = None  # the exception that made the density fail fatally
        self._logp = None
        self._cb = None
        self._expand_fn = None

    def _trampoline(self):
        from . import _lib

        if self._cb is not None:
            return self._cb
        logp = self._logp = self.make_logp_func()
        dim = self.dim

        def call(n, x, grad, out, _ud):  # PyDensity::logp, src/pyfunc.rs:206-230
            try:
                pos = np.ctypeslib.as_array(x, shape=(dim,)).copy()
                val = logp(pos)
                try:
                    lp, g = val
                    lp = float(lp)
                    g = np.asarray(g, dtype=np.float64)
                except Exception:
                    self.last_error = TypeError("logp function must return float.")
                    return -2  # ReturnTypeError: not recoverable
                if not np.isfinite(lp):
                    out[0] = lp
                    return 4  # BadLogp: recoverable
                if g.shape != (dim,):
                    self.last_error = ValueError("gradient has the wrong shape")
                    return -3
                C.memmove(grad, g.ctypes.data if g.flags["C_CONTIGUOUS"] else
                          np.ascontiguousarray(g).ctypes.data, 8 * dim)
                out[0] = lp
                return 0
            except BaseException as exc:  # PyError: recoverable iff exc.is_recoverable
                if getattr(exc, "is_recoverable", False):
                    return 1
                self.last_error = exc
                return -1

        self._cb = _lib.LOGP_FN(call)
        return self._cb

    def _descriptor(self):
        from . import _lib

        cb = self._trampoline()
        d = _lib.ModelDesc()
        d.kind = _lib.MODEL_KINDS["host"]
        d.dim = self.dim
        d.sigma = 1.0
        d.host_logp = C.cast(cb, C.c_void_p).value
        d.host_threads = int(self.host_threads)
        return d, [cb, self]

    def _expand(self, q):
        """PyDensity::expand_vector (src/pyfunc.rs:232-330): `expand(x)` returns a dict whose
        keys follow the variable table; shapes and dtypes are checked like the reference."""
        if self._expand_fn is None:
            self._expand_fn = self.make_expand_func(0, 0, 0)
        q = np.asarray(q, dtype=np.float64)
        lead = q.shape[:-1]
        flat = q.reshape(-1, self.dim)
        cols = {v.name: np.empty((len(flat),) + tuple(v.shape), dtype=v.dtype) for v in self.variables}
        for i, x in enumerate(flat):
            try:
                vals = self._expand_fn(np.array(x))
            except Exception as exc:
                raise RuntimeError("Expanding function raised an error") from exc
            if not isinstance(vals, dict):
                raise RuntimeError("Expand function did not return a dict")
            for v, (name, val) in zip(self.variables, vals.items()):
                if name != v.name:
                    raise RuntimeError(f"Unexpected expand key: expected {v.name} but found {name}")
                if val is None:
                    continue
                val = np.asarray(val)
                if val.dtype != v.dtype:
                    raise RuntimeError(f"variable {v.name} had incorrect type")
                if val.ndim != len(v.shape):
                    raise RuntimeError(f"unexpected number of dimensions for variable {v.name}")
                if tuple(val.shape) != tuple(v.shape):
                    raise RuntimeError(f"unexpected shape for variable {v.name}")
                cols[v.name][i] = val
        return {k: a.reshape(lead + a.shape[1:]) for k, a in cols.items()}


class StanLibrary:  # src/stan.rs:38-46
    def __init__(self, path):
        raise NotImplementedError("BridgeStan models need bridgestan/stanc, which this engine "
                                  "does not ship (SURVEY.md §2 N12)")


class StanModel:  # src/stan.rs:252-350
    def __init__(self, *a, **k):
        raise NotImplementedError("BridgeStan models need bridgestan/stanc, which this engine "
                                  "does not ship (SURVEY.md §2 N12)")


class _Store:
    """`_lib.store.*` (pyo3_object_store, src/wrapper.rs:1754-1756): only referenced in a type
    alias at import time (python/nutpie/sample.py:472-478); Zarr storage is out of scope."""

    def __init__(self, *a, **k):
        raise NotImplementedError("zarr/object-store storage is not supported by the B200 engine")


def _make_store_module():
    import types

    m = types.ModuleType("nutpie_b200._lib.store")
    for name in ("LocalStore", "S3Store", "GCSStore", "AzureStore", "HTTPStore", "MemoryStore"):
        setattr(m, name, type(name, (_Store,), {}))
    sys.modules.setdefault("nutpie_b200._lib.store", m)
    return m


store = _make_store_module()
