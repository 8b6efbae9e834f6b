"""Phase timing of one low-rank refresh (csrc/lowrank.cuh, lr_update) by ONE warp through the
component entry point, on a library built with -DNB200_LR_PROFILE (scripts/build_variant.py lrprof
nb200_api -DNB200_LR_PROFILE; NB200_LIB=nutpie_b200/variants/libnutpie_b200_lrprof.so)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nutpie_b200 import _lib

for dim, n in ((24, 100), (64, 160), (175, 20), (175, 160)):
    rng = np.random.default_rng(dim)
    a = rng.normal(size=(dim, dim))
    cov = a @ a.T / dim + 0.05 * np.eye(dim)
    x = rng.multivariate_normal(np.zeros(dim), cov, size=n)
    g = -np.linalg.solve(cov, x.T).T + 0.1 * rng.normal(size=(n, dim))
    _lib.lowrank_component(x, g)  # warm-up
    t = time.time()
    out = _lib.lowrank_component(x, g, cutoff=2.0, max_rank=32)
    print(f"dim {dim} n {n}: rank {len(out['vals'])}, wall {1e3 * (time.time() - t):.1f} ms", flush=True)
