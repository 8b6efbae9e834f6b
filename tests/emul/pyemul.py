"""ctypes loader for the host emulation of the CUDA sampler core (TEST INFRASTRUCTURE)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None
_LIB_NOBARRIER = None


class ModelDesc(C.Structure):
    """Mirror of nb200_model_desc (include/nutpie_b200.h)."""

    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("dim", C.c_uint64),
                ("mu", C.c_double), ("sigma", C.c_double),
                ("n_obs", C.c_int32), ("n_county", C.c_int32),
                ("y", C.c_void_p), ("county", C.c_void_p), ("floor", C.c_void_p)]


KINDS = {"normal": 1, "funnel": 2, "radon": 3}


def make_desc(kind: str, dim: int, **kw):
    d = ModelDesc()
    d.kind = KINDS[kind]
    d.dim = dim
    d.mu, d.sigma = float(kw.get("mu", 0.0)), float(kw.get("sigma", 1.0))
    keep = []
    if kind == "radon":
        y = np.ascontiguousarray(kw["y"], dtype=np.float64)
        county = np.ascontiguousarray(kw["county"], dtype=np.int32)
        floor = np.ascontiguousarray(kw["floor"], dtype=np.uint8)
        keep = [y, county, floor]
        d.n_obs, d.n_county = len(y), int(kw["n_county"])
        d.y, d.county, d.floor = y.ctypes.data, county.ctypes.data, floor.ctypes.data
    return d, keep


def lib():
    global _LIB
    if _LIB is None:
        subprocess.run(["make", "-C", str(_HERE), "libnuts_emul.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        _LIB = C.CDLL(str(_HERE / "libnuts_emul.so"))
        _LIB.emul_sample.restype = C.c_int
    return _LIB


def lib_nobarrier():
    """The core with ONE barrier of the leapfrog removed (negative control of the race check)."""
    global _LIB_NOBARRIER
    if _LIB_NOBARRIER is None:
        subprocess.run(["make", "-C", str(_HERE), "libnuts_emul_nobarrier.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        _LIB_NOBARRIER = C.CDLL(str(_HERE / "libnuts_emul_nobarrier.so"))
    return _LIB_NOBARRIER


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def sample(kind, dim, settings, n_chains, chain_id_offset=0, q0=None, init_mean=None,
           z_tape=None, max_per_launch=0, smem_slots=0, **model_kw):
    L = lib()
    desc, keep = make_desc(kind, dim, **model_kw)
    n_total = settings.num_tune + settings.num_draws
    n_rows = n_total if settings.save_warmup else settings.num_draws
    sdim = settings.store_dims if 0 < settings.store_dims < dim else dim
    # the engine's trace is row-major ([row][chain][...]); hand back chain-major views
    draws = np.zeros((n_rows, n_chains, sdim))
    stats = np.zeros((n_rows, n_chains, 16))
    grads = np.zeros((n_rows, n_chains, sdim)) if settings.store_gradient else None
    mm = np.zeros((n_rows, n_chains, sdim)) if settings.store_mass_matrix else None
    if q0 is not None:
        q0 = np.ascontiguousarray(q0, dtype=np.float64)
    if init_mean is not None:
        init_mean = np.ascontiguousarray(init_mean, dtype=np.float64)
    if z_tape is not None:
        z_tape = np.ascontiguousarray(z_tape, dtype=np.float64)
    steps = C.c_uint64(0)
    rc = L.emul_sample(C.byref(settings), C.byref(desc), C.c_uint64(n_chains),
                       C.c_uint64(chain_id_offset), _ptr(q0), _ptr(init_mean), _ptr(z_tape),
                       _ptr(draws), _ptr(stats), _ptr(grads), _ptr(mm), C.byref(steps),
                       C.c_int(max_per_launch), C.c_int(smem_slots))
    if rc != 0:
        raise RuntimeError(f"emul_sample failed: {rc}")
    tv = lambda a: None if a is None else a.transpose(1, 0, 2)
    return dict(draws=tv(draws), stats=tv(stats), gradients=tv(grads), mass_matrix_inv=tv(mm),
                total_steps=int(steps.value))


def sample_lanes(kind, dim, settings, n_chains, threads_per_chain=32, chain_id_offset=0, max_per_launch=0,
                 smem_slots=0, lane_order=0, drop_barrier=False, expand=False, **model_kw):
    """The same core run by `threads_per_chain` cooperative lanes per chain (GroupLanes in
    emul.cpp): the geometry, per-thread loops, shared-memory tier and density layouts the GPU uses.
    lane_order: 0 lanes run in ascending order between barriers, 1 descending, >= 2 a fresh
    pseudo-random order after every barrier."""
    L = lib_nobarrier() if drop_barrier else lib()
    L.emul_sample_lanes.restype = C.c_int
    desc, keep = make_desc(kind, dim, **model_kw)
    n_total = settings.num_tune + settings.num_draws
    n_rows = n_total if settings.save_warmup else settings.num_draws
    L.emul_expanded_dim.restype = C.c_uint64
    width = int(L.emul_expanded_dim(C.byref(desc))) if expand else dim
    draws = np.zeros((n_rows, n_chains, width))
    stats = np.zeros((n_rows, n_chains, 16))
    steps = C.c_uint64(0)
    rc = L.emul_sample_lanes(C.byref(settings), C.byref(desc), C.c_int(threads_per_chain), C.c_uint64(n_chains),
                             C.c_uint64(chain_id_offset), _ptr(draws), _ptr(stats), C.byref(steps),
                             C.c_int(max_per_launch), C.c_int(smem_slots), C.c_int(lane_order), C.c_int(1 if expand else 0))
    if rc != 0:
        raise RuntimeError(f"emul_sample_lanes failed: {rc}")
    return dict(draws=draws.transpose(1, 0, 2), stats=stats.transpose(1, 0, 2), total_steps=int(steps.value))
