"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE — see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package nutpie_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

NSTAT = 16
STAT_NAMES = [
    "depth", "maxdepth_reached", "index_in_trajectory", "logp", "energy", "energy_error",
    "diverging", "step_size", "step_size_bar", "n_steps", "mean_tree_accept",
    "mean_tree_accept_sym", "tuning", "draw", "chain", "reserved",
]


class Settings(C.Structure):
    """Mirror of nb200_settings (include/nutpie_b200.h)."""

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


def default_settings(**kw) -> Settings:
    """nuts_rs DiagNutsSettings::default() as recalled in SURVEY.md Appendix A.1."""
    s = Settings()
    s.seed = 0
    s.num_tune, s.num_draws = 400, 1000
    s.maxdepth, s.mindepth = 10, 0
    s.check_turning = 1
    s.store_gradient = 0
    s.store_mass_matrix = 0
    s.use_grad_based_estimate = 1
    s.max_energy_error = 1000.0
    s.initial_step, s.target_accept = 0.1, 0.8
    s.max_step_size = float("inf")
    s.da_k, s.da_t0, s.da_gamma = 0.75, 10.0, 0.05
    s.step_size_method = 0
    s.fixed_step_size = 0.1
    s.early_window, s.step_size_window = 0.3, 0.15
    s.mass_matrix_switch_freq, s.early_mass_matrix_switch_freq = 80, 10
    s.mass_matrix_update_freq = 1
    s.init_kind, s.num_try_init, s.init_radius = 0, 10, 2.0
    s.store_dims = 0
    s.save_warmup = 1
    s.store_divergences, s.adaptation = 0, 0
    s.adam_learning_rate, s.step_size_jitter = 0.05, 0.0
    s.mass_matrix_eigval_cutoff, s.mass_matrix_gamma = 2.0, 1e-5
    s.mass_matrix_max_rank = 32
    for k, v in kw.items():
        if not hasattr(s, k):
            raise AttributeError(k)
        setattr(s, k, v)
    return s


class NormalData(C.Structure):
    _fields_ = [("mu", C.c_double), ("sigma", C.c_double)]


class RadonData(C.Structure):
    _fields_ = [("n_obs", C.c_int32), ("n_county", C.c_int32), ("y", C.c_void_p),
                ("county", C.c_void_p), ("floor", C.c_void_p)]


LOGP_FN = C.CFUNCTYPE(C.c_int, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double),
                      C.POINTER(C.c_double), C.c_void_p)


def build(force: bool = False) -> Path:
    so = _HERE / "liboracle.so"
    srcs = [_HERE / n for n in ("nuts_oracle.c", "models.c", "lowrank.c", "oracle.h", "philox.h")]
    srcs.append(_HERE.parent / "include" / "nutpie_b200.h")
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(_HERE), "liboracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = _HERE / "liboracle.so"
        if not so.exists():
            build()
        else:
            try:
                build()
            except Exception:
                pass  # GPU box without sources newer than the .so: use as is
        _LIB = C.CDLL(str(so))
        _LIB.oracle_sample.restype = C.c_int
        _LIB.oracle_sample_ex.restype = C.c_int
        _LIB.oracle_leapfrog.restype = C.c_int
        _LIB.oracle_is_turning.restype = C.c_int
    return _LIB


_LIB_FAST = None
FAST_BUILD_KIND = None  # "fast" | "literal" (fallback): which build lib_fast() handed out


def _cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def lib_fast() -> C.CDLL:
    """The same sources built for SPEED on the machine that runs them (-O3 -march=native, FMA
    contraction allowed) — used only where the oracle is TIMED as the CPU baseline (bench.py), so
    that the baseline is not handicapped by the literal, unfused arithmetic the parity tests want.
    Rebuilt when the CPU model changes (the .so travels between machines)."""
    global _LIB_FAST, FAST_BUILD_KIND
    if _LIB_FAST is None:
        so, tag = _HERE / "liboracle_fast.so", _HERE / "liboracle_fast.cpu"
        cpu = _cpu_model()
        srcs = [_HERE / n for n in ("nuts_oracle.c", "models.c", "lowrank.c", "oracle.h", "philox.h")]
        stale = (not so.exists() or not tag.exists() or tag.read_text() != cpu
                 or any(x.stat().st_mtime > so.stat().st_mtime for x in srcs))
        try:
            if stale:
                subprocess.run(["make", "-B", "-C", str(_HERE), "liboracle_fast.so"], check=True,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
                tag.write_text(cpu)
            _LIB_FAST = C.CDLL(str(so))
            _LIB_FAST.oracle_sample.restype = C.c_int
            _LIB_FAST.oracle_sample_ex.restype = C.c_int
            FAST_BUILD_KIND = "fast"
        except (OSError, subprocess.CalledProcessError) as exc:
            import warnings

            warnings.warn(f"speed build of the oracle failed ({exc}); timing the literal build "
                          "(about 20% slower: the CPU baseline is understated)")
            _LIB_FAST = lib()  # no compiler on this machine: time the literal build
            FAST_BUILD_KIND = "literal"

    return _LIB_FAST


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Model:
    """A host density in the reference plug-in ABI plus its user_data."""

    def __init__(self, kind: str, dim: int, fast: bool = False, **kw):
        L = lib_fast() if fast else lib()  # fast: density from the speed build (timing only)
        self.kind, self.dim = kind, int(dim)
        self._keep = []
        if kind == "normal":
            self.fn = L.oracle_logp_normal
            self.ud = NormalData(float(kw.get("mu", 0.0)), float(kw.get("sigma", 1.0)))
        elif kind == "funnel":
            self.fn = L.oracle_logp_funnel
            self.ud = NormalData(0.0, 1.0)
        elif kind == "halfnormal":  # HalfNormal(1) on the log scale (the reference's golden model)
            self.fn = L.oracle_logp_halfnormal
            self.ud = NormalData(0.0, 1.0)
        elif kind == "radon":
            y = np.ascontiguousarray(kw["y"], dtype=np.float64)
            county = np.ascontiguousarray(kw["county"], dtype=np.int32)
            floor = np.ascontiguousarray(kw["floor"], dtype=np.uint8)
            self._keep += [y, county, floor]
            self.fn = L.oracle_logp_radon
            self.ud = RadonData(len(y), int(kw["n_county"]), _ptr(y), _ptr(county), _ptr(floor))
            assert self.dim == 2 * int(kw["n_county"]) + 5
        elif kind == "logreg":  # user_data = flat doubles [N, D, X, y]
            flat = np.ascontiguousarray(kw["data"], dtype=np.float64)
            self._keep.append(flat)
            self.fn = L.oracle_logp_logreg
            self.ud = None
            assert self.dim == int(flat[1])
        else:
            raise ValueError(kind)
        self.fn_ptr = C.cast(self.fn, C.c_void_p)
        if kind == "logreg":
            self.ud_ptr = C.c_void_p(self._keep[0].ctypes.data)
        else:
            self.ud_ptr = C.cast(C.pointer(self.ud), C.c_void_p)

    def logp_grad(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        single = q.ndim == 1
        q2 = q.reshape(-1, self.dim)
        g = np.empty_like(q2)
        lp = np.empty(len(q2))
        rc = np.empty(len(q2), dtype=np.int32)
        fn = C.cast(self.fn_ptr, LOGP_FN)
        for i in range(len(q2)):
            out = C.c_double()
            rc[i] = fn(self.dim, q2[i].ctypes.data_as(C.POINTER(C.c_double)),
                       g[i].ctypes.data_as(C.POINTER(C.c_double)), C.byref(out), self.ud_ptr)
            lp[i] = out.value
        if single:
            return lp[0], g[0], rc[0]
        return lp, g, rc


def sample(model: Model, settings: Settings, n_chains: int, chain_id_offset: int = 0,
           n_threads: int = 0, q0=None, init_mean=None, z_tape=None, fast: bool = False):
    """Run the oracle sampler.  Returns dict(draws, stats, gradients, mass_matrix_inv,
    total_steps) with the nb200_trace_view layout.  fast: use the speed build (timing only)."""
    L = lib_fast() if fast else lib()
    n_total = settings.num_tune + settings.num_draws
    n_rows = n_total if settings.save_warmup else settings.num_draws
    sdim = settings.store_dims if 0 < settings.store_dims < model.dim else model.dim
    draws = np.zeros((n_chains, n_rows, sdim))
    stats = np.zeros((n_chains, n_rows, NSTAT))
    grads = np.zeros((n_chains, n_rows, sdim)) if settings.store_gradient else None
    mm = np.zeros((n_chains, n_rows, sdim)) if settings.store_mass_matrix else None
    divs = np.zeros((n_chains, n_rows, 4, sdim)) if settings.store_divergences else None
    if q0 is not None:
        q0 = np.ascontiguousarray(q0, dtype=np.float64).reshape(n_chains, model.dim)
    if init_mean is not None:
        init_mean = np.ascontiguousarray(init_mean, dtype=np.float64).reshape(model.dim)
    if z_tape is not None:
        z_tape = np.ascontiguousarray(z_tape, dtype=np.float64).reshape(n_chains, n_total, model.dim)
    steps = C.c_uint64(0)
    low_rank = settings.adaptation == 1
    eig = (np.full((n_chains, n_rows, settings.mass_matrix_max_rank), np.nan)
           if settings.store_mass_matrix and low_rank else None)
    L.oracle_sample_lr.restype = C.c_int
    rc = L.oracle_sample_lr(C.byref(settings), model.fn_ptr, model.ud_ptr, C.c_uint64(model.dim),
                            C.c_uint64(n_chains), C.c_uint64(chain_id_offset), C.c_int(n_threads),
                            _ptr(q0), _ptr(init_mean), _ptr(z_tape), _ptr(draws), _ptr(stats),
                            _ptr(grads), _ptr(mm), _ptr(divs), _ptr(eig), C.byref(steps))
    if rc != 0:
        raise RuntimeError(f"oracle_sample failed with code {rc}")
    # low rank: the mass-matrix rows are mass_matrix_stds, eigvals the kept eigenvalues
    return dict(draws=draws, stats=stats, gradients=grads, mass_matrix_inv=mm,
                mass_matrix_eigvals=eig, divergences=divs, total_steps=int(steps.value))


def lowrank_update(draws, grads, gamma=1e-5, cutoff=2.0, max_rank=32, stds0=None):
    """oracle_lowrank_update on a window [n][dim]: returns (stds, vals [k], vecs [k][dim])."""
    L = lib()
    draws = np.ascontiguousarray(draws, dtype=np.float64)
    grads = np.ascontiguousarray(grads, dtype=np.float64)
    n, dim = draws.shape
    max_rank = min(max_rank, dim)
    stds = np.ones(dim) if stds0 is None else np.array(stds0, dtype=np.float64)
    vals, vecs = np.zeros(max_rank + 1), np.zeros((max_rank + 1, dim))
    k = C.c_size_t(0)
    L.oracle_lowrank_update.restype = C.c_int
    rc = L.oracle_lowrank_update(C.c_size_t(dim), C.c_size_t(n), _ptr(draws), _ptr(grads),
                                 C.c_double(gamma), C.c_double(cutoff), C.c_size_t(max_rank),
                                 _ptr(stds), _ptr(vals), _ptr(vecs), C.byref(k))
    if rc != 0:
        raise RuntimeError(f"oracle_lowrank_update failed with code {rc}")
    return stds, vals[:k.value].copy(), vecs[:k.value].copy()


def lowrank_velocity(stds, vals, vecs, p):
    L = lib()
    L.oracle_lowrank_velocity.restype = None
    p = np.ascontiguousarray(p, dtype=np.float64)
    v = np.empty_like(p)
    vecs = np.ascontiguousarray(vecs, dtype=np.float64)
    L.oracle_lowrank_velocity(C.c_size_t(len(p)), _ptr(np.ascontiguousarray(stds)), C.c_size_t(len(vals)),
                              _ptr(np.ascontiguousarray(vals)), _ptr(vecs), _ptr(p), _ptr(v))
    return v


def lowrank_momentum(stds, vals, vecs, z):
    L = lib()
    L.oracle_lowrank_momentum.restype = None
    z = np.ascontiguousarray(z, dtype=np.float64)
    p = np.empty_like(z)
    vecs = np.ascontiguousarray(vecs, dtype=np.float64)
    L.oracle_lowrank_momentum(C.c_size_t(len(z)), _ptr(np.ascontiguousarray(stds)), C.c_size_t(len(vals)),
                              _ptr(np.ascontiguousarray(vals)), _ptr(vecs), _ptr(z), _ptr(p))
    return p


def stat(stats: np.ndarray, name: str) -> np.ndarray:
    return stats[..., STAT_NAMES.index(name)]
