"""Compiled device models — the B200 counterpart of nutpie's CompiledModel.

In the reference a CompiledModel wraps a HOST logp function pointer produced by
numba / BridgeStan / a Python callable (python/nutpie/compile_pymc.py:104-236,
compile_stan.py:17-130, compiled_pyfunc.py:14-155) and hands it to
`_lib.PySampler.from_pymc/from_stan/from_pyfunc` in `_make_sampler`
(compile_pymc.py:168-187).  The B200 engine evaluates the density on the
device, so a compiled model here is a descriptor of one of the hand-written
CUDA densities (nutpie_b200/csrc/models.cuh) plus the host-side `expand`
(CpuLogpFunc::expand_vector, src/pymc.rs:217-286) that turns unconstrained
draws into named, constrained variables.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib


@dataclass
class CompiledDeviceModel:
    kind: str
    n_dim: int
    dims: dict = field(default_factory=dict)      # variable name -> tuple of dim names
    coords: dict = field(default_factory=dict)    # dim name -> coordinate values
    shapes: dict = field(default_factory=dict)    # variable name -> shape
    params: dict = field(default_factory=dict)

    # -- what _lib.PySampler needs ------------------------------------------
    def _descriptor(self):
        d = _lib.ModelDesc()
        d.kind = _lib.MODEL_KINDS[self.kind]
        d.dim = self.n_dim
        d.mu = float(self.params.get("mu", 0.0))
        d.sigma = float(self.params.get("sigma", 1.0))
        keep = []
        if self.kind == "radon":
            y = np.ascontiguousarray(self.params["y"], dtype=np.float64)
            county = np.ascontiguousarray(self.params["county"], dtype=np.int32)
            floor = np.ascontiguousarray(self.params["floor"], dtype=np.uint8)
            keep = [y, county, floor]
            d.n_obs, d.n_county = len(y), int(self.params["n_county"])
            d.y, d.county, d.floor = y.ctypes.data, county.ctypes.data, floor.ctypes.data
        if self.kind == "custom":
            src = self.params["cuda_source"].encode("utf-8")
            data = np.ascontiguousarray(self.params.get("data", ()), dtype=np.float64).ravel()
            keep = [src, data]
            d.cuda_source = src
            d.user_data = data.ctypes.data if data.size else None
            d.n_user_data = data.size
            d.n_user_scratch = int(self.params.get("scratch", 0))
        self._keep = keep
        return d, keep

    def _variable_dims(self):
        return dict(self.dims)

    def _expand(self, q: np.ndarray) -> dict:
        """unconstrained draws [..., n_dim] -> {variable: constrained values}"""
        q = np.asarray(q)
        if self.kind == "normal":
            return {"x": q if self.n_dim > 1 else q[..., 0]}
        if self.kind == "funnel":
            return {"log_sigma": q[..., 0], "x": q[..., 1:]}
        if self.kind == "custom":
            return self._split_expanded(q)
        if self.kind == "radon":
            J = int(self.params["n_county"])
            sd_a, sd_b = np.exp(q[..., J + 1]), np.exp(q[..., 2 * J + 3])
            raw_a, raw_b = q[..., 1:J + 1], q[..., J + 3:2 * J + 3]
            return {
                "intercept": q[..., 0],
                "county_raw": raw_a,
                "county_sd": sd_a,
                "floor_effect": q[..., J + 2],
                "county_floor_raw": raw_b,
                "county_floor_sd": sd_b,
                "sigma": np.exp(q[..., 2 * J + 4]),
                "county_effect": raw_a * sd_a[..., None],
                "county_floor_effect": raw_b * sd_b[..., None],
            }
        raise ValueError(self.kind)

    def expanded_layout(self):
        """[(name, start, end, shape)] over the device-side expanded vector — the role of
        PyVariable.start_idx/end_idx (src/common.rs:283-300)."""
        if self.kind == "normal":
            return [("x", 0, self.n_dim, (self.n_dim,) if self.n_dim > 1 else ())]
        if self.kind == "funnel":
            return [("log_sigma", 0, 1, ()), ("x", 1, self.n_dim, (self.n_dim - 1,))]
        if self.kind == "custom":
            out, pos = [], 0
            for name, shape in self.shapes.items():
                n = int(np.prod(shape)) if shape else 1
                out.append((name, pos, pos + n, tuple(shape)))
                pos += n
            return out
        if self.kind == "radon":
            J, D = int(self.params["n_county"]), self.n_dim
            return [("intercept", 0, 1, ()), ("county_raw", 1, J + 1, (J,)),
                    ("county_sd", J + 1, J + 2, ()), ("floor_effect", J + 2, J + 3, ()),
                    ("county_floor_raw", J + 3, 2 * J + 3, (J,)),
                    ("county_floor_sd", 2 * J + 3, 2 * J + 4, ()), ("sigma", 2 * J + 4, D, ()),
                    ("county_effect", D, D + J, (J,)), ("county_floor_effect", D + J, D + 2 * J, (J,))]
        raise ValueError(self.kind)

    def _split_expanded(self, e: np.ndarray) -> dict:
        """expanded draws [..., n_expanded] (computed on the device) -> {variable: values}"""
        out = {}
        for name, a, b, shape in self.expanded_layout():
            if not shape:
                out[name] = e[..., a]
            elif len(shape) == 1:
                out[name] = e[..., a:b]
            else:
                out[name] = e[..., a:b].reshape(e.shape[:-1] + tuple(shape))
        return out

    # names of value variables that are stored transformed (compile_pymc.py:810-814)
    @property
    def reparameterized_names(self):
        return ["county_sd", "county_floor_sd", "sigma"] if self.kind == "radon" else []

    def _make_sampler(self, settings, init_mean, cores, progress_type, extra_callback=None,
                      extra_callback_rate=500, store=None, **kw):
        """compile_pymc.py:168-187 — build the model object and start the sampler."""
        return _lib.PySampler.from_device_model(settings, cores, self, progress_type,
                                                extra_callback, extra_callback_rate, store,
                                                init_mean=init_mean, **kw)

    def with_data(self, **updates):
        """compile_pymc.py:136-161 — replace data arrays, shapes must match."""
        params = dict(self.params)
        for k, v in updates.items():
            if k not in params:
                raise KeyError(f"Unknown shared variable: {k}")
            v = np.asarray(v)
            if np.shape(params[k]) != v.shape:
                raise RuntimeError(f"Shared variable {k} has the wrong shape")
            params[k] = v
        return CompiledDeviceModel(self.kind, self.n_dim, dict(self.dims), dict(self.coords),
                                   dict(self.shapes), params)


def normal_model(dim: int = 1, mu: float = 0.0, sigma: float = 1.0) -> CompiledDeviceModel:
    """x ~ Normal(mu, sigma) iid over `dim` coordinates (README.md:148-163 Stan example;
    BASELINE.json configs 1 and 4)."""
    dims = {"x": ("x_dim",)} if dim > 1 else {"x": ()}
    return CompiledDeviceModel("normal", int(dim), dims, {}, {"x": (dim,) if dim > 1 else ()},
                               dict(mu=mu, sigma=sigma))


def funnel_model(dim: int = 9) -> CompiledDeviceModel:
    """Neal's funnel (docs/sample-stats.qmd:19-21; BASELINE.json config 5)."""
    return CompiledDeviceModel("funnel", int(dim), {"log_sigma": (), "x": ("x_dim",)}, {},
                               {"log_sigma": (), "x": (dim - 1,)}, {})


def radon_model(y, county, floor, n_county: int, county_names=None) -> CompiledDeviceModel:
    """The hierarchical radon model of README.md:53-88 with plain-Normal raw effects
    (notebooks/pytensor_logp.md:57-88); n_dim = 2 * n_county + 5."""
    J = int(n_county)
    coords = {"county": np.asarray(county_names) if county_names is not None else np.arange(J)}
    dims = {
        "intercept": (), "county_raw": ("county",), "county_sd": (), "floor_effect": (),
        "county_floor_raw": ("county",), "county_floor_sd": (), "sigma": (),
        "county_effect": ("county",), "county_floor_effect": ("county",),
    }
    shapes = {k: ((J,) if v else ()) for k, v in dims.items()}
    return CompiledDeviceModel("radon", 2 * J + 5, dims, coords, shapes,
                               dict(y=np.asarray(y, dtype=np.float64),
                                    county=np.asarray(county, dtype=np.int32),
                                    floor=np.asarray(floor, dtype=np.uint8), n_county=J))


def custom_model(ndim: int, cuda_source: str, data=None, *, scratch: int = 0, shapes=None,
                 dims=None, coords=None) -> CompiledDeviceModel:
    """A density given as CUDA source (NB200_MODEL_CUSTOM, include/nutpie_b200.h): the device
    counterpart of `nutpie.compiled_pyfunc.from_pyfunc` (python/nutpie/compiled_pyfunc.py:108-155),
    where the user hands over a function instead of a PyMC / Stan model.  `cuda_source` defines

        __device__ int nb200_user_logp(const nb200_group& grp, int dim, const double* q,
                                       double* grad, double* logp_partial, const double* data);

    `data` is a flat float64 array the density reads on the device, `scratch` the number of
    doubles of per-chain shared memory it gets as `grp.scratch`.  `shapes` maps variable
    names to shapes that partition the `ndim` unconstrained coordinates in order (default: one
    vector "x"); custom densities have no transforms, so the trace holds q itself."""
    ndim = int(ndim)
    if shapes is None:
        shapes = {"x": (ndim,)}
    shapes = {k: tuple(int(n) for n in v) for k, v in shapes.items()}
    total = sum(int(np.prod(v)) if v else 1 for v in shapes.values())
    if total != ndim:
        raise ValueError(f"variable shapes cover {total} coordinates, the density has {ndim}")
    if dims is None:
        dims = {k: tuple(f"{k}_dim_{i}" for i in range(len(v))) for k, v in shapes.items()}
    params = dict(cuda_source=str(cuda_source), scratch=int(scratch),
                  data=np.ascontiguousarray(() if data is None else data, dtype=np.float64).ravel())
    return CompiledDeviceModel("custom", ndim, dict(dims), dict(coords or {}), shapes, params)
