"""Host-side mirror of `nutpie._lib` over the C-ABI of libnutpie_b200.so.

The reference's `_lib` is a PyO3 module (src/wrapper.rs:1738-1758) that owns a
`nuts_rs::Sampler`.  There is no Rust toolchain in this image, so the same
Python-visible surface — PyNutsSettings, PySampler, PyChainProgress, PyTrace,
ProgressType, PyStorage — is written in Python on top of `ctypes` bindings of
include/nutpie_b200.h; every class cites the PyO3 item it mirrors.  All
sampling work happens in the CUDA library: if it cannot be loaded this module
raises, there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import threading
import time
from pathlib import Path

import numpy as np

__version__ = "0.1.0"

_PKG = Path(__file__).resolve().parent
_SO = _PKG / "libnutpie_b200.so"
if os.environ.get("NB200_LIB"):  # a build variant of the same sources (scripts/build_variants.py)
    _SO = Path(os.environ["NB200_LIB"]).resolve()
_lib = None
_lib_lock = threading.Lock()

NSTAT = 16
STAT_NAMES = [
    "depth", "maxdepth_reached", "index_in_trajectory", "logp", "energy", "energy_error",
    "diverging", "step_size", "step_size_bar", "n_steps", "mean_tree_accept",
    "mean_tree_accept_sym", "tuning", "draw", "chain", "reserved",
]
_STAT_DTYPES = {
    "depth": np.uint64, "maxdepth_reached": np.bool_, "index_in_trajectory": np.int64,
    "diverging": np.bool_, "n_steps": np.uint64, "tuning": np.bool_, "draw": np.uint64,
    "chain": np.uint64,
}

# python/nutpie/sample.py:641-646
DIVERGENCE_COLUMNS = ["divergence_start", "divergence_end", "divergence_momentum",
                      "divergence_start_gradient"]
NB200_ETIMEOUT = 1
_ERRORS = {-1: ValueError, -2: RuntimeError, -3: ValueError, -4: RuntimeError, -5: RuntimeError,
           -6: RuntimeError}


class Settings(C.Structure):
    """nb200_settings (include/nutpie_b200.h)."""

    _fields_ = [
        ("seed", C.c_uint64), ("num_tune", C.c_uint64), ("num_draws", C.c_uint64),
        ("maxdepth", C.c_uint32), ("mindepth", C.c_uint32),
        ("check_turning", C.c_int32), ("store_gradient", C.c_int32),
        ("store_mass_matrix", C.c_int32), ("use_grad_based_estimate", C.c_int32),
        ("max_energy_error", C.c_double),
        ("initial_step", C.c_double), ("target_accept", C.c_double),
        ("max_step_size", C.c_double), ("da_k", C.c_double), ("da_t0", C.c_double),
        ("da_gamma", C.c_double),
        ("step_size_method", C.c_int32), ("_pad0", C.c_int32),
        ("fixed_step_size", C.c_double),
        ("early_window", C.c_double), ("step_size_window", C.c_double),
        ("mass_matrix_switch_freq", C.c_uint64),
        ("early_mass_matrix_switch_freq", C.c_uint64),
        ("mass_matrix_update_freq", C.c_uint64),
        ("init_kind", C.c_int32), ("num_try_init", C.c_int32),
        ("init_radius", C.c_double),
        ("store_dims", C.c_uint64),
        ("save_warmup", C.c_int32), ("expand_draws", C.c_int32),
        ("store_divergences", C.c_int32), ("adaptation", C.c_int32),
        ("adam_learning_rate", C.c_double), ("step_size_jitter", C.c_double),
        ("mass_matrix_eigval_cutoff", C.c_double), ("mass_matrix_gamma", C.c_double),
        ("mass_matrix_max_rank", C.c_uint64),
    ]


# the reference's plug-in ABI (src/pymc.rs:23-37; numba side compile_pymc.py:975-981, 1018-1024)
LOGP_FN = C.CFUNCTYPE(C.c_int, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double),
                      C.POINTER(C.c_double), C.c_void_p)
EXPAND_FN = C.CFUNCTYPE(C.c_int, C.c_size_t, C.c_size_t, C.POINTER(C.c_double),
                        C.POINTER(C.c_double), C.c_void_p)


class ModelDesc(C.Structure):
    """nb200_model_desc."""

    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("dim", C.c_uint64),
                ("mu", C.c_double), ("sigma", C.c_double),
                ("n_obs", C.c_int32), ("n_county", C.c_int32),
                ("y", C.c_void_p), ("county", C.c_void_p), ("floor", C.c_void_p),
                ("cuda_source", C.c_char_p), ("user_data", C.c_void_p),
                ("n_user_data", C.c_uint64), ("n_user_scratch", C.c_uint64),
                ("host_logp", C.c_void_p), ("host_user_data", C.c_void_p),
                ("host_expand", C.c_void_p), ("host_expand_user_data", C.c_void_p),
                ("host_expanded_dim", C.c_uint64),
                ("host_threads", C.c_int32), ("_pad2", C.c_int32)]


class Progress(C.Structure):
    """nb200_progress."""

    _fields_ = [("finished_draws", C.c_uint64), ("total_draws", C.c_uint64),
                ("divergences", C.c_uint64), ("latest_num_steps", C.c_uint64),
                ("total_num_steps", C.c_uint64), ("step_size", C.c_double),
                ("tuning", C.c_int32), ("started", C.c_int32)]


MODEL_KINDS = {"normal": 1, "funnel": 2, "radon": 3, "custom": 4, "host": 5}
ABI_VERSION = 4
LOW_RANK_SUPPORTED = True  # csrc/lowrank.cuh (built-in and host plug-in densities)


def library_path() -> Path:
    return _SO


def load_library() -> C.CDLL:
    """Load libnutpie_b200.so, building it with nvcc when sources are newer.
    Raises RuntimeError when the CUDA library is unavailable — never falls back."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if os.environ.get("NUTPIE_B200_NO_BUILD") != "1" and not os.environ.get("NB200_LIB"):
            try:
                from . import build as _build

                if _build.needs_build():
                    _build.build()
            except Exception as exc:  # no nvcc on this host: use the shipped .so if any
                if not _SO.exists():
                    raise RuntimeError(
                        f"libnutpie_b200.so is missing and could not be built ({exc}); "
                        "the B200 engine has no CPU fallback") from exc
        if not _SO.exists():
            raise RuntimeError("libnutpie_b200.so is missing; run `python -m nutpie_b200.build`")
        L = C.CDLL(str(_SO))
        L.nb200_abi_version.restype = C.c_int
        L.nb200_last_error.restype = C.c_char_p
        L.nb200_device_count.restype = C.c_int
        L.nb200_model_expanded_dim.restype = C.c_uint64
        L.nb200_model_expanded_dim.argtypes = [C.POINTER(ModelDesc)]
        L.nb200_sampler_create.restype = C.c_void_p
        L.nb200_sampler_create.argtypes = [C.POINTER(Settings), C.POINTER(ModelDesc), C.c_uint64,
                                           C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        for name in ("start", "is_finished", "pause", "resume", "abort", "destroy"):
            f = getattr(L, f"nb200_sampler_{name}")
            f.restype = C.c_int
            f.argtypes = [C.c_void_p]
        L.nb200_sampler_wait.restype = C.c_int
        L.nb200_sampler_wait.argtypes = [C.c_void_p, C.c_double]
        L.nb200_sampler_progress.restype = C.c_int
        L.nb200_sampler_progress.argtypes = [C.c_void_p, C.POINTER(Progress)]
        L.nb200_sampler_trace_into.restype = C.c_int
        L.nb200_sampler_trace_into.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.nb200_sampler_kernel_ms.restype = C.c_double
        L.nb200_sampler_kernel_ms.argtypes = [C.c_void_p]
        L.nb200_sampler_launch_count.restype = C.c_uint64
        L.nb200_sampler_launch_count.argtypes = [C.c_void_p]
        L.nb200_sampler_geometry.restype = C.c_int
        L.nb200_sampler_geometry.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 3
        L.nb200_sampler_device_buffers.restype = C.c_int
        L.nb200_sampler_device_buffers.argtypes = [C.c_void_p, C.POINTER(C.c_void_p),
                                                   C.POINTER(C.c_void_p)]
        L.nb200_sampler_set_draws_per_launch.restype = C.c_int
        L.nb200_sampler_set_draws_per_launch.argtypes = [C.c_void_p, C.c_uint64]
        L.nb200_sampler_set_trace_target.restype = C.c_int
        L.nb200_sampler_set_trace_target.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                                     C.c_size_t]
        L.nb200_sampler_trace_bytes.restype = C.c_int
        L.nb200_sampler_trace_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_size_t),
                                                C.POINTER(C.c_size_t)]
        L.nb200_sampler_divergence_trace_into.restype = C.c_int
        L.nb200_sampler_divergence_trace_into.argtypes = [C.c_void_p, C.c_void_p]
        L.nb200_sampler_set_trace_target_strided.restype = C.c_int
        L.nb200_sampler_set_trace_target_strided.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                                             C.c_size_t]
        L.nb200_sampler_eigvals_trace_into.restype = C.c_int
        L.nb200_sampler_eigvals_trace_into.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.nb200_lowrank_component.restype = C.c_int
        L.nb200_lowrank_component.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                              C.c_double, C.c_double, C.c_uint64, C.c_uint64] + \
            [C.c_void_p] * 7 + [C.POINTER(C.c_uint64)]
        L.nb200_host_expand_rows.restype = C.c_int
        L.nb200_host_expand_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t,
                                             C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        L.nb200_sampler_set_z_tape.restype = C.c_int
        L.nb200_sampler_set_z_tape.argtypes = [C.c_void_p, C.c_void_p]
        L.nb200_settings_default.restype = None
        L.nb200_settings_default.argtypes = [C.POINTER(Settings)]
        L.nb200_set_threads_per_chain.restype = None
        L.nb200_set_threads_per_chain.argtypes = [C.c_int32]
        L.nb200_set_chains_per_block.restype = None
        L.nb200_set_chains_per_block.argtypes = [C.c_int32]
        L.nb200_set_smem_slots.restype = None
        L.nb200_set_smem_slots.argtypes = [C.c_int32]
        L.nb200_set_stage_loads.restype = None
        L.nb200_set_stage_loads.argtypes = [C.c_int32]
        L.nb200_set_unroll.restype = None
        L.nb200_set_unroll.argtypes = [C.c_int32]
        L.nb200_set_pipeline.restype = None
        L.nb200_set_pipeline.argtypes = [C.c_int32]
        L.nb200_sampler_is_pipelined.restype = C.c_int
        L.nb200_sampler_is_pipelined.argtypes = [C.c_void_p]
        L.nb200_sampler_smem.restype = C.c_int
        L.nb200_sampler_smem.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.nb200_host_alloc.restype = C.c_void_p
        L.nb200_host_alloc.argtypes = [C.c_size_t]
        L.nb200_host_free.restype = None
        L.nb200_host_free.argtypes = [C.c_void_p]
        L.nb200_logp_grad.restype = C.c_int
        L.nb200_logp_grad.argtypes = [C.POINTER(ModelDesc), C.c_int, C.c_uint64] + [C.c_void_p] * 4
        L.nb200_leapfrog.restype = C.c_int
        L.nb200_leapfrog.argtypes = [C.POINTER(ModelDesc), C.c_int, C.c_uint64] + [C.c_void_p] * 15
        L.nb200_custom_model_compile.restype = C.c_int
        L.nb200_custom_model_compile.argtypes = [C.POINTER(ModelDesc), C.c_int, C.c_int,
                                                 C.c_char_p, C.c_size_t]
        if L.nb200_abi_version() != ABI_VERSION:
            raise RuntimeError("libnutpie_b200.so ABI version mismatch")
        _lib = L
        return L


def _check(rc: int):
    if rc >= 0:
        return rc
    msg = load_library().nb200_last_error().decode("utf-8", "replace")
    raise _ERRORS.get(rc, RuntimeError)(msg or f"nb200 error {rc}")


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class PinnedArray:
    """numpy array backed by pinned host memory from nb200_host_alloc."""

    def __init__(self, shape, dtype=np.float64):
        self._L = load_library()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = self._L.nb200_host_alloc(max(n, 8))
        if not self._p:
            raise MemoryError("nb200_host_alloc failed")
        buf = (C.c_char * max(n, 8)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            if self._p:
                self.array = None
                self._L.nb200_host_free(self._p)
                self._p = None
        except Exception:
            pass


# --------------------------------------------------------------------------
# settings façade — PyNutsSettings (src/wrapper.rs:106-826)
# --------------------------------------------------------------------------
_FLAT_DIRECT = {
    "maxdepth": ("maxdepth", int), "mindepth": ("mindepth", int),
    "check_turning": ("check_turning", bool),
    "initial_step": ("initial_step", float), "target_accept": ("target_accept", float),
    "max_step_size": ("max_step_size", float),
    "store_mass_matrix": ("store_mass_matrix", bool),
    "use_grad_based_mass_matrix": ("use_grad_based_estimate", bool),
    "mass_matrix_switch_freq": ("mass_matrix_switch_freq", int),
    "window_switch_freq": ("mass_matrix_switch_freq", int),
    "early_window_switch_freq": ("early_mass_matrix_switch_freq", int),
    "store_gradient": ("store_gradient", bool),
    "num_tune": ("num_tune", int), "num_draws": ("num_draws", int),
    "max_energy_error": ("max_energy_error", float),
    "store_divergences": ("store_divergences", bool),
    "step_size_adam_learning_rate": ("adam_learning_rate", float),
    "mass_matrix_eigval_cutoff": ("mass_matrix_eigval_cutoff", float),
    "mass_matrix_gamma": ("mass_matrix_gamma", float),
    "mass_matrix_max_rank": ("mass_matrix_max_rank", int),
    # ours (not in the reference): initial-point and trace controls
    "init_radius": ("init_radius", float), "num_try_init": ("num_try_init", int),
    "store_dims": ("store_dims", int),
}
# options that exist in the reference but belong to samplers / adaptations that
# are out of scope for the B200 engine
_UNSUPPORTED = {
    "target_integration_time": None, "extra_doublings": 0, "train_on_orbit": None,
    "store_transformed": False,
    "microcanonical_trajectory": False, "exact_normal_trajectory": False,
}
# options that only exist for one adaptation (wrapper.rs:138-145: ValueError otherwise)
_LOW_RANK_ONLY = ("mass_matrix_eigval_cutoff", "mass_matrix_gamma", "mass_matrix_max_rank")
_DIAG_ONLY = ("use_grad_based_mass_matrix",)


class PyNutsSettings:
    """Mirror of PyNutsSettings (src/wrapper.rs:106-110, 525-620, 716-770)."""

    def __init__(self, kind: str, seed=None):
        object.__setattr__(self, "_kind", kind)
        s = Settings()
        load_library().nb200_settings_default(C.byref(s))
        if seed is None:  # random_seed(), src/wrapper.rs:453-458
            seed = int.from_bytes(os.urandom(8), "little")
        s.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        object.__setattr__(self, "_c", s)
        object.__setattr__(self, "_num_chains", 6)
        object.__setattr__(self, "_store_unconstrained", False)

    # static constructors: src/wrapper.rs:718-737
    @staticmethod
    def Diag(seed=None):
        return PyNutsSettings("diag", seed)

    @staticmethod
    def LowRank(seed=None):
        # nuts-rs LowRankNutsSettings::default() [recalled]: the metric is refreshed every 10
        # draws (a refresh is an eigen-decomposition, not a vector pass), 800 tuning draws
        s = PyNutsSettings("low_rank", seed)
        s._c.adaptation = 1
        s._c.mass_matrix_update_freq = 10
        s._c.num_tune = 800
        return s

    @staticmethod
    def Flow(seed=None):
        raise NotImplementedError(
            "adaptation='flow' is not supported by the B200 engine (diag / draw_diag only)")

    def update(self, updates: dict):  # src/wrapper.rs:739-745
        for k, v in updates.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):  # apply_update, src/wrapper.rs:563-620
        c = self._c
        if name == "num_chains":
            object.__setattr__(self, "_num_chains", int(value))
        elif name == "store_unconstrained":
            object.__setattr__(self, "_store_unconstrained", bool(value))
        elif name == "step_size_adapt_method":
            if not isinstance(value, str):
                raise ValueError("step_size_adapt_method must be a string")
            if value == "dual_average":
                c.step_size_method = 0
            elif value == "adam":
                c.step_size_method = 1
            else:
                try:
                    c.fixed_step_size = float(value)
                except ValueError:
                    raise ValueError("step_size_adapt_method must be a positive float when "
                                     "using fixed step size") from None
                c.step_size_method = 2
        elif name == "step_size_jitter":  # Option<f64>, wrapper.rs:393-407
            c.step_size_jitter = 0.0 if value is None else float(value)
        elif name in _FLAT_DIRECT:
            if (name in _LOW_RANK_ONLY and self._kind != "low_rank") or \
                    (name in _DIAG_ONLY and self._kind != "diag"):
                raise ValueError(f"Option {name} not available for {self._kind} adaptation")
            field, typ = _FLAT_DIRECT[name]
            setattr(c, field, typ(value))
        elif name in _UNSUPPORTED:
            if value not in (None, False, 0, _UNSUPPORTED[name]):
                raise ValueError(f"Option {name} not available for the B200 diag engine")
        else:
            raise AttributeError(f"Unknown settings attribute: {name}")

    @property
    def num_chains(self):
        return self._num_chains

    @property
    def num_tune(self):
        return int(self._c.num_tune)

    @property
    def num_draws(self):
        return int(self._c.num_draws)

    @property
    def seed(self):
        return int(self._c.seed)

    def as_dict(self):  # src/wrapper.rs:751-769
        c = self._c
        flat = {name: getattr(c, name) for name, _ in Settings._fields_ if not name.startswith("_")}
        flat["num_chains"] = self._num_chains
        return {"sampler": "nuts", "adaptation": self._kind, "settings": flat}

    def update_settings(self, nested: dict):  # src/wrapper.rs:747-749
        for k, v in nested.get("settings", nested).items():
            if k == "num_chains":
                self.num_chains = v
            elif hasattr(self._c, k):
                setattr(self._c, k, v)
            else:
                raise AttributeError(f"Unknown settings attribute: {k}")

    def _copy_c(self) -> Settings:
        s = Settings()
        C.memmove(C.byref(s), C.byref(self._c), C.sizeof(Settings))
        return s


class PyMclmcSettings:
    """src/wrapper.rs:772-826 — the MCLMC sampler is out of scope for the B200 engine."""

    @staticmethod
    def _no(*a, **k):
        raise NotImplementedError("sampler='mclmc' is not supported by the B200 engine")

    Diag = LowRank = Flow = _no


class PyChainProgress:
    """Mirror of PyChainProgress (src/wrapper.rs:38-104)."""

    def __init__(self, p: Progress, runtime_ms: float, divergent_draws):
        self.finished_draws = int(p.finished_draws)
        self.total_draws = int(p.total_draws)
        self.divergences = int(p.divergences)
        self.started = bool(p.started)
        self.tuning = bool(p.tuning)
        self.latest_num_steps = int(p.latest_num_steps)
        self.num_steps = int(p.latest_num_steps)
        self.total_num_steps = int(p.total_num_steps)
        self.step_size = float(p.step_size)
        self.runtime_ms = float(runtime_ms)
        self.divergent_draws = list(divergent_draws)


class ProgressType:
    """Mirror of ProgressType (src/wrapper.rs:907-930).  Terminal/HTML rendering
    (src/progress.rs) is out of scope; callbacks receive PyChainProgress lists."""

    def __init__(self, kind, rate_ms=500, callback=None):
        self.kind, self.rate_ms, self.callback = kind, rate_ms, callback

    @staticmethod
    def none():
        return ProgressType("none")

    @staticmethod
    def indicatif(rate_ms):
        return ProgressType("indicatif", rate_ms)

    @staticmethod
    def template_callback(rate_ms, template, n_cores, callback):
        return ProgressType("template_callback", rate_ms, callback)


class PyStorage:
    """Mirror of PyStorage (src/wrapper.rs:940-951).  Zarr stores are out of scope."""

    def __init__(self, kind):
        self.kind = kind

    @staticmethod
    def arrow():
        return PyStorage("arrow")

    @staticmethod
    def zarr(store):
        raise NotImplementedError("zarr storage is not supported by the B200 engine")


class PyTrace:
    """Mirror of PyTrace (src/wrapper.rs:1467-1494): the trace of all chains.

    `draws` / `stats` are [chain, row, ...] views of the engine's row-major buffers.

    Besides the reference's `get_arrow_trace()` (one (posterior, sample_stats)
    RecordBatch pair per chain) it exposes the raw arrays (`draws`, `stats`),
    which is what the end-to-end path uses to avoid a per-chain Arrow hop."""

    def __init__(self, draws, stats, rows_filled, gradients=None, mass_matrix_inv=None,
                 variables=None, expand=None, keep=None, expanded=False, divergences=None,
                 mass_matrix_eigvals=None, low_rank=False):
        self.draws, self.stats, self.rows_filled = draws, stats, rows_filled
        # adaptation="low_rank" + store_mass_matrix: the mass-matrix rows are mass_matrix_stds and
        # mass_matrix_eigvals [chain, row, max_rank] holds the eigenvalues in use (NaN-padded)
        self.mass_matrix_eigvals, self.low_rank = mass_matrix_eigvals, low_rank
        # store_divergences: [chain, row, 4, dim] = start location, end location, start
        # momentum, start gradient of the diverging leapfrog (NaN rows otherwise)
        self.divergences = divergences
        self.expanded = expanded  # draws hold expanded vectors (constrained + deterministics)
        self.expand_fn = expand
        self.gradients, self.mass_matrix_inv = gradients, mass_matrix_inv
        self.variables, self._expand, self._keep = variables, expand, keep
        self._taken = False

    def mass_matrix_columns(self):
        """The store_mass_matrix columns under the reference's names (python/nutpie/sample.py:
        631-650): mass_matrix_inv (diag) | mass_matrix_stds + mass_matrix_eigvals (low rank)."""
        if self.low_rank:
            return [("mass_matrix_stds", self.mass_matrix_inv),
                    ("mass_matrix_eigvals", self.mass_matrix_eigvals)]
        return [("mass_matrix_inv", self.mass_matrix_inv)]

    def is_zarr(self):
        return False

    def is_arrow(self):
        return True

    def stat(self, name):
        a = self.stats[..., STAT_NAMES.index(name)]
        dt = _STAT_DTYPES.get(name)
        return a.astype(dt) if dt is not None else a

    def get_arrow_trace(self):
        """(posterior RecordBatches, sample_stats RecordBatches), one of each per chain —
        `Vec<ArrowTrace>` unzipped exactly as src/wrapper.rs:1477-1494 returns it; single-take
        like the reference."""
        import pyarrow as pa

        if self._taken:
            raise ValueError("The trace was already taken")
        self._taken = True
        out_draws, out_stats = [], []
        n_chains = self.draws.shape[0]
        for c in range(n_chains):
            n = int(self.rows_filled[c])
            q = self.draws[c, :n]
            cols, fields = [], []
            values = self._expand(q) if self._expand is not None else {"unconstrained_draw": q}
            for name, arr in values.items():
                arr = np.ascontiguousarray(arr)
                shape = arr.shape[1:]
                size = int(np.prod(shape)) if shape else 1
                meta = {"dims": "", "shape": ",".join(map(str, shape))}
                if self.variables and name in self.variables:
                    meta["dims"] = ",".join(self.variables[name])
                if shape:
                    col = pa.FixedSizeListArray.from_arrays(pa.array(arr.reshape(-1)), size)
                else:
                    col = pa.array(arr)
                cols.append(col)
                fields.append(pa.field(name, col.type, metadata=meta))
            posterior = pa.RecordBatch.from_arrays(cols, schema=pa.schema(fields))
            scols, sfields = [], []
            for name in STAT_NAMES[:-1]:
                a = self.stats[c, :n, STAT_NAMES.index(name)]
                dt = _STAT_DTYPES.get(name)
                col = pa.array(a.astype(dt) if dt is not None else a)
                scols.append(col)
                sfields.append(pa.field(name, col.type, metadata={"dims": "", "shape": ""}))
            vec_stats = [("gradient", self.gradients)] + self.mass_matrix_columns()
            if self.divergences is not None:
                for k, name in enumerate(DIVERGENCE_COLUMNS):
                    vec_stats.append((name, self.divergences[:, :, k]))
            for name, arr in vec_stats:
                if arr is not None:
                    a = np.ascontiguousarray(arr[c, :n])
                    col = pa.FixedSizeListArray.from_arrays(pa.array(a.reshape(-1)), a.shape[1])
                    if name in DIVERGENCE_COLUMNS:  # null unless the draw diverged
                        mask = np.isnan(a).all(axis=1)
                        col = pa.FixedSizeListArray.from_arrays(pa.array(a.reshape(-1)), a.shape[1],
                                                                mask=pa.array(mask))
                    scols.append(col)
                    dim_name = "mass_matrix_eigvals_dim" if name == "mass_matrix_eigvals" \
                        else "unconstrained_parameter"
                    sfields.append(pa.field(name, col.type, metadata={
                        "dims": dim_name, "shape": str(a.shape[1])}))
            stats = pa.RecordBatch.from_arrays(scols, schema=pa.schema(sfields))
            out_draws.append(posterior)
            out_stats.append(stats)
        return out_draws, out_stats


class PySampler:
    """Mirror of PySampler (src/wrapper.rs:953-1457) driving one nb200_sampler.

    `from_device_model` takes the place of from_pymc / from_stan / from_pyfunc
    (src/wrapper.rs:1187-1250): the model is a device density descriptor."""

    def __init__(self, settings: PyNutsSettings, model, *, n_chains=None, chain_id_offset=0,
                 device=0, progress_type=None, init_mean=None, q0=None, z_tape=None,
                 draws_per_launch=0, autostart=True, trace_buffers=None):
        L = load_library()
        self._L = L
        self._settings = settings
        self._model = model
        self._c = settings._copy_c()
        self.n_chains = int(n_chains if n_chains is not None else settings.num_chains)
        self._desc, self._keep = model._descriptor()
        self.dim = int(self._desc.dim)
        self._host_model = int(self._desc.kind) == MODEL_KINDS["host"]
        if init_mean is not None:
            init_mean = np.ascontiguousarray(init_mean, dtype=np.float64).reshape(self.dim)
        if q0 is None and self._host_model:
            # Model::init_position (src/pymc.rs:505-534, src/pyfunc.rs:535-569): the model's
            # own `init_func(seed)`; without one the engine draws U(-2, 2) on the device
            q0 = model._initial_points(settings.seed, self.n_chains, int(chain_id_offset))
        if q0 is not None:
            q0 = np.ascontiguousarray(q0, dtype=np.float64).reshape(self.n_chains, self.dim)
        h = L.nb200_sampler_create(C.byref(self._c), C.byref(self._desc), self.n_chains,
                                   int(chain_id_offset), int(device), _ptr(q0), _ptr(init_mean))
        if not h:
            msg = L.nb200_last_error().decode("utf-8", "replace")
            raise (ValueError if "must" in msg or "unknown" in msg else RuntimeError)(msg)
        self._h = C.c_void_p(h)
        self._lock = threading.RLock()  # guards _h against close() racing the progress thread
        self.n_total = int(self._c.num_tune + self._c.num_draws)
        self.n_rows = self.n_total if self._c.save_warmup else int(self._c.num_draws)
        sd = int(self._c.store_dims)
        self.grad_dim = sd if 0 < sd < self.dim else self.dim
        self.expanded = bool(self._c.expand_draws) and not (0 < sd < self.dim) and not self._host_model
        self.sdim = int(L.nb200_model_expanded_dim(C.byref(self._desc))) if self.expanded else self.grad_dim
        try:
            if z_tape is not None:
                z_tape = np.ascontiguousarray(z_tape, dtype=np.float64)
                _check(L.nb200_sampler_set_z_tape(self._h, _ptr(z_tape)))
            if draws_per_launch:
                _check(L.nb200_sampler_set_draws_per_launch(self._h, int(draws_per_launch)))
            if trace_buffers is False:  # no streaming: one copy when the trace is taken
                trace_buffers = None
            elif trace_buffers is None:
                # the default call: plain (pageable) result arrays, registered as the streaming
                # target all the same — finished rows land in them while the kernel runs (through
                # the engine's pinned staging ring, nb200_api.cu d2h_block), so that the trace is
                # on the host when sampling ends instead of one serial copy afterwards
                sh = self.trace_shapes()
                trace_buffers = {"draws": np.empty(sh["draws"]), "stats": np.empty(sh["stats"])}
            self._trace_buffers = None
            if trace_buffers is not None:
                self._register_trace_buffers(trace_buffers)
        except Exception:
            L.nb200_sampler_destroy(self._h)
            self._h = None
            raise
        self._taken = False
        self._t_start = None
        self._progress_type = progress_type or ProgressType.none()
        self._progress_thread = None
        self._stop_progress = threading.Event()
        self._started = False
        if autostart:
            self.start()

    def _register_trace_buffers(self, trace_buffers):
        """Rows are streamed into these host arrays while sampling runs (before start only)."""
        L = self._L
        sh = self.trace_shapes()
        strided = False
        for k in ("draws", "stats"):
            a = trace_buffers[k]
            if tuple(a.shape) != sh[k]:  # checked BEFORE the engine may write into them
                raise ValueError(f"trace buffer '{k}' must be a row-major array of shape "
                                 f"{sh[k]}, got {tuple(a.shape)}")
            # dense, or the column block [:, a:b, :] of a wider row-major array (the
            # shards of a multi-GPU job share ONE result array): rows may be strided
            w = a.shape[2]
            ok = (a.dtype == np.float64 and a.strides[2] == 8 and a.strides[1] == 8 * w
                  and a.strides[0] >= 8 * w * a.shape[1] and a.strides[0] % 8 == 0)
            if not ok:
                raise ValueError("trace buffers must be float64 arrays, C-contiguous or a "
                                 "block of chains of a C-contiguous [row][chain][width] array")
            strided |= not a.flags["C_CONTIGUOUS"]
        if strided:
            _check(L.nb200_sampler_set_trace_target_strided(
                self._h, C.c_void_p(trace_buffers["draws"].ctypes.data),
                trace_buffers["draws"].strides[0] // 8,
                C.c_void_p(trace_buffers["stats"].ctypes.data), trace_buffers["stats"].strides[0] // 8))
        else:
            _check(L.nb200_sampler_set_trace_target(
                self._h, _ptr(trace_buffers["draws"]), trace_buffers["draws"].nbytes,
                _ptr(trace_buffers["stats"]), trace_buffers["stats"].nbytes))
        self._trace_buffers = trace_buffers

    def start(self):
        """Launch the sampling kernel (non-blocking), like nuts_rs::Sampler::new returning
        while its workers run (src/wrapper.rs:983-990)."""
        if self._started:
            raise ValueError("sampler already started")
        _check(self._L.nb200_sampler_start(self._h))
        self._started = True
        self._t_start = time.perf_counter()
        if self._progress_type.callback is not None:
            self._progress_thread = threading.Thread(target=self._progress_loop, daemon=True)
            self._progress_thread.start()

    @staticmethod
    def from_device_model(settings, cores, model, progress_type=None, extra_callback=None,
                          extra_callback_rate=None, store=None, **kw):
        pt = progress_type
        if extra_callback is not None and (pt is None or pt.callback is None):
            pt = ProgressType("callback", extra_callback_rate or 500, extra_callback)
        devices = kw.pop("devices", None)
        if devices is not None and not (isinstance(devices, (list, tuple)) and len(devices) == 1):
            kw.pop("device", None)
            return PyMultiSampler(settings, model, devices=devices, progress_type=pt, **kw)
        if devices is not None:
            kw["device"] = int(devices[0])
        return PySampler(settings, model, progress_type=pt, **kw)

    @staticmethod
    def from_pymc(settings, cores, model, progress_type=None, extra_callback=None,
                  extra_callback_rate=500, store=None, **kw):
        """src/wrapper.rs:1190-1208 — `model` is a PyMcModel (LogpFunc + ExpandFunc pointers):
        sampled through the host plug-in service, `cores` host threads calling the pointer."""
        if not isinstance(model, PyMcModel):
            raise TypeError("from_pymc needs a PyMcModel")
        kw.pop("init_mean", None)  # passed down but unused by _make_model (compile_pymc.py:189)
        model.host_threads = int(cores or 0)
        return PySampler.from_device_model(settings, cores, model, progress_type, extra_callback,
                                           extra_callback_rate, store, **kw)

    @staticmethod
    def from_pyfunc(settings, cores, model, progress_type=None, extra_callback=None,
                    extra_callback_rate=500, store=None, **kw):
        """src/wrapper.rs:1232-1250 — `model` is a PyModel (Python callables)."""
        if not isinstance(model, PyModel):
            raise TypeError("from_pyfunc needs a PyModel")
        kw.pop("init_mean", None)
        return PySampler.from_device_model(settings, cores, model, progress_type, extra_callback,
                                           extra_callback_rate, store, **kw)

    @staticmethod
    def from_stan(*a, **k):
        """src/wrapper.rs:1211-1229 — BridgeStan is out of scope."""
        raise NotImplementedError("BridgeStan models are not supported by the B200 engine")

    # -- progress ---------------------------------------------------------
    def progress(self):
        arr = (Progress * self.n_chains)()
        with self._lock:
            if self._h is None:
                raise ValueError("sampler is closed")
            _check(self._L.nb200_sampler_progress(self._h, arr))
        rt = (time.perf_counter() - self._t_start) * 1e3 if self._t_start else 0.0
        return [PyChainProgress(arr[i], rt, ()) for i in range(self.n_chains)]

    def _progress_loop(self):
        rate = max(self._progress_type.rate_ms, 10) / 1e3
        while not self._stop_progress.wait(rate):
            try:
                self._progress_type.callback(self.progress())
            except Exception as exc:  # src/progress.rs:426-431: printed, not raised
                import sys

                print(f"progress callback failed: {exc}", file=sys.stderr)
            with self._lock:
                if self._h is None or self._L.nb200_sampler_is_finished(self._h):
                    break

    # -- control (src/wrapper.rs:1252-1365) -------------------------------
    def wait(self, timeout_seconds=None):
        """Block until finished.  Raises TimeoutError like wrapper.rs:1113 and
        polls in 100 ms slices so KeyboardInterrupt is honoured (wrapper.rs:1099-1143)."""
        deadline = None if timeout_seconds is None else time.perf_counter() + timeout_seconds
        while True:
            slice_s = 0.1
            if deadline is not None:
                slice_s = min(slice_s, max(deadline - time.perf_counter(), 0.0))
            try:
                rc = _check(self._L.nb200_sampler_wait(self._h, slice_s))
            except RuntimeError as exc:
                cause = getattr(self._model, "last_error", None)
                if cause is not None:  # the Python density raised (src/pyfunc.rs:100-116)
                    raise RuntimeError(str(exc)) from cause
                raise
            if rc == 0:
                return
            if deadline is not None and time.perf_counter() >= deadline:
                raise TimeoutError("Timeout while waiting for sampler to finish")

    def pause(self):
        _check(self._L.nb200_sampler_pause(self._h))

    def resume(self):
        _check(self._L.nb200_sampler_resume(self._h))

    def abort(self):
        _check(self._L.nb200_sampler_abort(self._h))

    def is_finished(self):
        return bool(self._L.nb200_sampler_is_finished(self._h))

    def is_empty(self, ignore_error=False):
        return self._taken

    def flush(self):
        return None

    # -- results (src/wrapper.rs:1401-1456) --------------------------------
    def trace_shapes(self):
        """Shapes of the engine's ROW-major host buffers: [row][chain][width]."""
        return {"draws": (self.n_rows, self.n_chains, self.sdim),
                "stats": (self.n_rows, self.n_chains, NSTAT),
                "gradients": (self.n_rows, self.n_chains, self.grad_dim)}

    def _trace(self, out=None) -> PyTrace:
        sh = self.trace_shapes()
        if out is None:
            out = self._trace_buffers
        if out is not None:
            draws, stats = out["draws"], out["stats"]
            if draws.shape != sh["draws"] or stats.shape != sh["stats"]:
                raise ValueError(f"trace buffers must be row-major arrays of shape {sh['draws']} / {sh['stats']}")
        else:
            draws, stats = np.empty(sh["draws"]), np.empty(sh["stats"])
        grads = np.empty(sh["gradients"]) if self._c.store_gradient else None
        mm = np.empty(sh["gradients"]) if self._c.store_mass_matrix else None
        rows = np.zeros(self.n_chains, dtype=np.uint64)
        _check(self._L.nb200_sampler_trace_into(self._h, _ptr(draws), _ptr(stats), _ptr(grads),
                                                _ptr(mm), _ptr(rows)))
        divs = None
        if self._c.store_divergences:
            divs = np.empty((self.n_rows, self.n_chains, 4, self.grad_dim))
            _check(self._L.nb200_sampler_divergence_trace_into(self._h, _ptr(divs)))
            divs = divs.transpose(1, 0, 2, 3)
        eig = None
        low_rank = self._c.adaptation == 1
        if low_rank and self._c.store_mass_matrix:
            width = C.c_uint64(0)
            _check(self._L.nb200_sampler_eigvals_trace_into(self._h, None, C.byref(width)))
            eig = np.empty((self.n_rows, self.n_chains, int(width.value)))
            _check(self._L.nb200_sampler_eigvals_trace_into(self._h, _ptr(eig), None))
            eig = eig.transpose(1, 0, 2)
        if self.expanded:  # rows already hold the expanded vector: slice it into variables
            expand = self._model._split_expanded
        elif self.sdim == self.dim:
            expand = self._model._expand
        else:
            expand = None
        # chain-major VIEWS ([chain, row, ...]) of the row-major buffers
        tv = lambda a: None if a is None else a.transpose(1, 0, 2)
        return PyTrace(tv(draws), tv(stats), rows, tv(grads), tv(mm),
                       variables=self._model._variable_dims(), expand=expand,
                       keep=[draws, stats, grads, mm], expanded=self.expanded, divergences=divs,
                       mass_matrix_eigvals=eig, low_rank=low_rank)

    def inspect(self, out=None):
        return self._trace(out)

    def take_results(self, out=None):
        if not self.is_finished():
            raise ValueError("Sampler is still running")  # src/wrapper.rs:1431-1440
        if self._taken:
            raise ValueError("Sampler is empty")
        tr = self._trace(out)
        self._taken = True
        return tr

    # -- measurement hooks --------------------------------------------------
    def kernel_ms(self):
        return float(self._L.nb200_sampler_kernel_ms(self._h))

    def launch_count(self):
        return int(self._L.nb200_sampler_launch_count(self._h))

    def geometry(self):
        a, b, g = C.c_int32(), C.c_int32(), C.c_int32()
        _check(self._L.nb200_sampler_geometry(self._h, C.byref(a), C.byref(b), C.byref(g)))
        sl, by = C.c_int32(), C.c_int32()
        _check(self._L.nb200_sampler_smem(self._h, C.byref(sl), C.byref(by)))
        return dict(threads_per_chain=a.value, block=b.value, grid=g.value,
                    smem_slots=sl.value, smem_bytes_per_chain=by.value,
                    pipelined=bool(self._L.nb200_sampler_is_pipelined(self._h)))

    def device_buffers(self):
        d, s = C.c_void_p(), C.c_void_p()
        _check(self._L.nb200_sampler_device_buffers(self._h, C.byref(d), C.byref(s)))
        return d.value, s.value

    def close(self):
        self._stop_progress.set()
        t = self._progress_thread
        if t is not None and t is not threading.current_thread():
            t.join()  # the loop wakes on the event; a slow user callback is waited for, never
            #           raced: the native handle must outlive every call that uses it
        with self._lock:
            if self._h is not None:
                self._L.nb200_sampler_destroy(self._h)
                self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass



class MultiTrace:
    """The traces of the device shards of one run, presented as one PyTrace: chains are in global
    order (device 0's block first); `draws` / `stats` concatenate the per-device blocks on first
    access (each block is a view of that device's own — pinned, if supplied — host buffer)."""

    def __init__(self, parts, whole=None):
        self.parts = parts
        self.expanded = parts[0].expanded
        self.expand_fn = parts[0].expand_fn
        self.variables = parts[0].variables
        self._expand = parts[0]._expand
        self._cache = {}
        if whole is not None:  # the shards wrote into one [row][chain][width] array: chain-major views
            self._cache = {k: a.transpose(1, 0, 2) for k, a in whole.items()}
        self._taken = False

    def _cat(self, name):
        if name not in self._cache:
            vals = [getattr(p, name) for p in self.parts]
            self._cache[name] = None if vals[0] is None else np.concatenate(vals, axis=0)
        return self._cache[name]

    draws = property(lambda self: self._cat("draws"))
    stats = property(lambda self: self._cat("stats"))
    rows_filled = property(lambda self: self._cat("rows_filled"))
    gradients = property(lambda self: self._cat("gradients"))
    mass_matrix_inv = property(lambda self: self._cat("mass_matrix_inv"))
    divergences = property(lambda self: self._cat("divergences"))
    mass_matrix_eigvals = property(lambda self: self._cat("mass_matrix_eigvals"))
    low_rank = property(lambda self: self.parts[0].low_rank)
    mass_matrix_columns = PyTrace.mass_matrix_columns

    def is_zarr(self):
        return False

    def is_arrow(self):
        return True

    def stat(self, name):
        a = self.stats[..., STAT_NAMES.index(name)]
        dt = _STAT_DTYPES.get(name)
        return a.astype(dt) if dt is not None else a

    def get_arrow_trace(self):
        if self._taken:
            raise ValueError("The trace was already taken")
        self._taken = True
        d, s = [], []
        for p in self.parts:
            a, b = p.get_arrow_trace()
            d += a
            s += b
        return d, s


class PyMultiSampler:
    """One sampling job over several GPUs of ONE process — the analogue of `cores` in
    nuts_rs::Sampler::new (src/wrapper.rs:977-1085; python/nutpie/sample.py:1061-1075 picks it):
    chains are split into contiguous blocks of global chain ids, one PySampler (own stream,
    own persistent kernel, own trace buffers) per device, one host thread per device while
    waiting so that every device's finished rows stream to the host concurrently.  The random
    streams are keyed by global chain id, so the result equals the single-device run chain for
    chain (tests/test_gpu_parity.py::test_chain_sharding_reproduces_single_run)."""

    def __init__(self, settings, model, *, n_chains=None, devices=None, chain_id_offset=0,
                 progress_type=None, trace_buffers=None, q0=None, z_tape=None, **kw):
        from .distributed import shard

        n_chains = int(n_chains if n_chains is not None else settings.num_chains)
        if devices is None:
            devices = list(range(device_count()))
        elif isinstance(devices, int):
            devices = list(range(devices))
        devices = [int(d) for d in devices][:max(1, n_chains)]
        if not devices:
            raise RuntimeError("no CUDA device available: the B200 engine has no CPU fallback")
        self.devices = devices
        self.n_chains = n_chains
        self.parts = []
        self._offsets = []
        self._big = None
        self._progress_type = progress_type or ProgressType.none()
        self._model = model
        self._stop_progress = threading.Event()
        self._progress_thread = None
        try:
            for i, dev in enumerate(devices):
                n_local, off = shard(n_chains, i, len(devices))
                if n_local == 0:
                    continue
                sl = slice(off, off + n_local)
                bufs = False  # registered below: every shard writes into the job's ONE result array
                if trace_buffers is not None:  # per-device list of dict(draws=, stats=)
                    bufs = trace_buffers[i]
                self._offsets.append((off, n_local))
                self.parts.append(PySampler(
                    settings, model, n_chains=n_local, chain_id_offset=int(chain_id_offset) + off,
                    device=dev, trace_buffers=bufs, autostart=False,
                    q0=None if q0 is None else np.asarray(q0)[sl],
                    z_tape=None if z_tape is None else np.asarray(z_tape)[sl], **kw))
            if trace_buffers is None and self.parts:
                # one [row][chain][width] array for the whole job: device i streams its chains into
                # its own block of columns (row-strided targets), nothing is concatenated afterwards
                sh = self.parts[0].trace_shapes()
                self._big = {"draws": np.empty((sh["draws"][0], n_chains, sh["draws"][2])),
                             "stats": np.empty((sh["stats"][0], n_chains, sh["stats"][2]))}
                for p, (off, n_local) in zip(self.parts, self._offsets):
                    p._register_trace_buffers({k: a[:, off:off + n_local, :] for k, a in self._big.items()})
            for p in self.parts:  # all devices start before anyone waits
                p.start()
        except Exception:
            self.close()
            raise
        self._t_start = time.perf_counter()
        if self._progress_type.callback is not None:
            self._progress_thread = threading.Thread(target=self._progress_loop, daemon=True)
            self._progress_thread.start()

    def _progress_loop(self):
        rate = max(self._progress_type.rate_ms, 10) / 1e3
        while not self._stop_progress.wait(rate):
            try:
                self._progress_type.callback(self.progress())
            except Exception as exc:
                import sys

                print(f"progress callback failed: {exc}", file=sys.stderr)
            if self.is_finished():
                break

    def progress(self):
        return [c for p in self.parts for c in p.progress()]

    def _each(self, fn):
        """fn(part) on one host thread per device; re-raises the first failure."""
        errs = [None] * len(self.parts)

        def run(i, p):
            try:
                fn(p)
            except BaseException as exc:  # noqa: BLE001 - re-raised below
                errs[i] = exc

        ths = [threading.Thread(target=run, args=(i, p)) for i, p in enumerate(self.parts)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        for e in errs:
            if e is not None:
                raise e

    def wait(self, timeout_seconds=None):
        self._each(lambda p: p.wait(timeout_seconds))

    def pause(self):
        self._each(lambda p: p.pause())

    def resume(self):
        self._each(lambda p: p.resume())

    def abort(self):
        self._each(lambda p: p.abort())

    def is_finished(self):
        return all(p.is_finished() for p in self.parts)

    def is_empty(self, ignore_error=False):
        return all(p.is_empty(ignore_error) for p in self.parts)

    def flush(self):
        return None

    def inspect(self, out=None):
        return MultiTrace([p.inspect() for p in self.parts], whole=self._big)

    def take_results(self, out=None):
        if not self.is_finished():
            raise ValueError("Sampler is still running")
        res = [None] * len(self.parts)

        def take(i, p):
            res[i] = p.take_results()

        ths = [threading.Thread(target=take, args=(i, p)) for i, p in enumerate(self.parts)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if any(r is None for r in res):
            raise ValueError("Sampler is empty")
        return MultiTrace(res, whole=self._big)

    def kernel_ms(self):
        """device time of the slowest device (the job's time)"""
        return max(p.kernel_ms() for p in self.parts)

    def launch_count(self):
        return sum(p.launch_count() for p in self.parts)

    def geometry(self):
        g = self.parts[0].geometry()
        g["devices"] = list(self.devices)
        return g

    def close(self):
        self._stop_progress.set()
        t = self._progress_thread
        if t is not None and t is not threading.current_thread():
            t.join()
        for p in self.parts:
            p.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def PySamplerDeferred(*args, **kwargs):
    """A PySampler whose device state is allocated and initialised but whose kernel is
    only launched by .start() — lets a benchmark time the resident-input path."""
    return PySampler(*args, autostart=False, **kwargs)


# --------------------------------------------------------------------------
# component entry points (parity tests)
# --------------------------------------------------------------------------
def compile_custom(model, threads_per_chain=32, dims_per_thread=0):
    """NVRTC-compile a custom CUDA density for one kernel geometry (no GPU needed);
    raises RuntimeError carrying the compiler log on failure."""
    L = load_library()
    desc, keep = model._descriptor()
    log = C.create_string_buffer(1 << 16)
    rc = L.nb200_custom_model_compile(C.byref(desc), threads_per_chain, dims_per_thread, log,
                                      len(log))
    if rc != 0:
        _check(rc)
    return True


def logp_grad(model, q, device=0):
    L = load_library()
    desc, keep = model._descriptor()
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, int(desc.dim))
    n = len(q)
    lp, g, rc = np.empty(n), np.empty_like(q), np.empty(n, dtype=np.int32)
    _check(L.nb200_logp_grad(C.byref(desc), device, n, _ptr(q), _ptr(lp), _ptr(g), _ptr(rc)))
    return lp, g, rc


def leapfrog(model, q, p, g, var, p_sum, eps, direction, idx, device=0):
    L = load_library()
    desc, keep = model._descriptor()
    D = int(desc.dim)
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64).reshape(-1, D)
    q, p, g, var, p_sum = map(f, (q, p, g, var, p_sum))
    n = len(q)
    eps = np.ascontiguousarray(np.broadcast_to(eps, (n,)), dtype=np.float64)
    direction = np.ascontiguousarray(np.broadcast_to(direction, (n,)), dtype=np.int32)
    idx = np.ascontiguousarray(np.broadcast_to(idx, (n,)), dtype=np.int64)
    qo, po, go, so = (np.empty_like(q) for _ in range(4))
    lp, kin, rc = np.empty(n), np.empty(n), np.empty(n, dtype=np.int32)
    _check(L.nb200_leapfrog(C.byref(desc), device, n, _ptr(q), _ptr(p), _ptr(g), _ptr(var),
                            _ptr(p_sum), _ptr(eps), _ptr(direction), _ptr(idx), _ptr(qo),
                            _ptr(po), _ptr(go), _ptr(so), _ptr(lp), _ptr(kin), _ptr(rc)))
    return dict(q=qo, p=po, g=go, p_sum=so, logp=lp, kinetic=kin, rc=rc)


def lowrank_component(draws, grads, gamma=1e-5, cutoff=2.0, max_rank=32, p=None, z=None, device=0):
    """Refresh the low-rank metric on the device from a window [n][dim] of draws / gradients and
    apply it: returns dict(stds, vals [k], vecs [k][dim], velocity = M^-1 p, momentum = M^1/2 z)."""
    L = load_library()
    draws = np.ascontiguousarray(draws, dtype=np.float64)
    grads = np.ascontiguousarray(grads, dtype=np.float64)
    n, dim = draws.shape
    max_rank = min(int(max_rank), dim)
    f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64).reshape(-1, dim)
    p, z = f(p), f(z)
    n_vec = max(len(p) if p is not None else 0, len(z) if z is not None else 0)
    if p is None and n_vec:
        p = np.zeros((n_vec, dim))
    if z is None and n_vec:
        z = np.zeros((n_vec, dim))
    v, pm = np.zeros((n_vec, dim)), np.zeros((n_vec, dim))
    stds, vals, vecs = np.zeros(dim), np.zeros(max_rank), np.zeros((max_rank, dim))
    k = C.c_uint64(0)
    _check(L.nb200_lowrank_component(device, dim, n, _ptr(draws), _ptr(grads), gamma, cutoff, max_rank,
                                     n_vec, _ptr(p), _ptr(v), _ptr(z), _ptr(pm), _ptr(stds), _ptr(vals),
                                     _ptr(vecs), C.byref(k)))
    kk = int(k.value)
    return dict(stds=stds, vals=vals[:kk].copy(), vecs=vecs[:kk].copy(), velocity=v, momentum=pm)


def set_threads_per_chain(t: int):
    load_library().nb200_set_threads_per_chain(int(t))


def set_chains_per_block(c: int):
    load_library().nb200_set_chains_per_block(int(c))


def set_smem_slots(n: int):
    load_library().nb200_set_smem_slots(int(n))


def set_unroll(on: bool):
    load_library().nb200_set_unroll(1 if on else 0)


def set_pipeline(on: bool):
    """Two warps per chain (integrator + tree) for gathering densities: True = auto, False = off."""
    load_library().nb200_set_pipeline(1 if on else 0)


def set_stage_loads(mode):
    """Streaming-leapfrog mode bits: 1 bulk-copy staging, 2 alternating sweep direction,
    4 L2 eviction hints (True = 1, False = 0)."""
    load_library().nb200_set_stage_loads(int(mode))


def device_count() -> int:
    return int(load_library().nb200_device_count())


# model-side classes of `nutpie._lib` (src/pymc.rs, src/pyfunc.rs, src/common.rs, src/stan.rs)
from ._lib_models import (ExpandFunc, LogpFunc, PyMcModel, PyModel, PyVariable,  # noqa: E402
                          StanLibrary, StanModel, chain_seed, store)
