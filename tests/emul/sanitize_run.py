"""Runs the serial and the lane emulation of the sampler core from a library built with
-fsanitize=address,undefined (TEST INFRASTRUCTURE; started by tests/test_core_emulation.py in a
subprocess with libasan preloaded).  Any out-of-bounds index into the pool, the shared-memory tier,
the front buffer or the density scratch — or undefined behaviour in the bookkeeping — aborts it."""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from tests.emul import pyemul as E  # noqa: E402

E._LIB = C.CDLL(sys.argv[1])
E._LIB.emul_sample.restype = C.c_int

from nutpie_b200.datasets import make_radon_data  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

d = make_radon_data()
J = d["n_county"]
D = 2 * J + 5
kw = dict(y=d["y"], county=d["county"], floor=d["floor"], n_county=J)
s = O.default_settings(seed=3, num_tune=16, num_draws=4, init_radius=1.0)
for T, slots in ((32, 3), (64, 0), (32, 40)):
    b = E.sample_lanes("radon", D, s, 1, threads_per_chain=T, smem_slots=slots, lane_order=2, max_per_launch=13, **kw)
    print("radon lanes", T, slots, b["total_steps"])
s2 = O.default_settings(seed=5, num_tune=25, num_draws=10, maxdepth=12)
for kind, dim, T in (("funnel", 9, 32), ("normal", 37, 32), ("normal", 100, 32), ("normal", 40, 64)):
    b = E.sample_lanes(kind, dim, s2, 2, threads_per_chain=T, smem_slots=5)
    print(kind, dim, T, b["total_steps"])
for kind, dim in (("normal", 33), ("funnel", 9), ("normal", 1)):
    b = E.sample(kind, dim, s2, 2, smem_slots=64, max_per_launch=7)
    print("serial", kind, dim, b["total_steps"])
b = E.sample("radon", D, s, 1, smem_slots=4, **kw)
print("serial radon", b["total_steps"])
# low-rank engine: whole runs (full-space and subspace refreshes, window deque, fifth slot vector)
# and the refresh alone at sizes where the subspace scratch is tight
import numpy as np  # noqa: E402

slr = O.default_settings(seed=11, num_tune=60, num_draws=10, adaptation=1, store_mass_matrix=1,
                         mass_matrix_update_freq=10)
rng = np.random.default_rng(0)
small = dict(y=rng.normal(1.0, 0.8, 60), county=rng.integers(0, 4, 60).astype(np.int32),
             floor=rng.integers(0, 2, 60).astype(np.uint8), n_county=4)
b = E.sample("radon", 13, slr, 2, **small)
print("serial low-rank radon-13", b["total_steps"])
b = E.sample("normal", 40, slr, 2, mu=1.0, sigma=2.0)
print("serial low-rank normal-40 (subspace refreshes)", b["total_steps"])
b = E.sample_lanes("normal", 40, slr, 1, threads_per_chain=32, lane_order=2, mu=1.0, sigma=2.0)
print("lanes low-rank normal-40", b["total_steps"])
E._LIB.emul_lowrank_component.restype = C.c_int
for dim, n in ((13, 11), (16, 4), (17, 5), (40, 8), (50, 15), (64, 19), (175, 52), (9, 3), (1, 5)):
    x, g = rng.normal(size=(n, dim)), rng.normal(size=(n, dim))
    p = np.eye(dim)
    v, pm = np.zeros_like(p), np.zeros_like(p)
    stds, vals, vecs = np.zeros(dim), np.zeros(dim), np.zeros((dim, dim))
    k = C.c_uint64(0)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = E._LIB.emul_lowrank_component(C.c_uint64(dim), C.c_uint64(n), ptr(x), ptr(g), C.c_double(1e-5),
                                       C.c_double(2.0), C.c_uint64(dim), C.c_uint64(dim), ptr(p), ptr(v), ptr(p),
                                       ptr(pm), ptr(stds), ptr(vals), ptr(vecs), C.byref(k))
    print("low-rank refresh", dim, n, "rc", rc, "rank", k.value)
print("SANITIZE-OK")
