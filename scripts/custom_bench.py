"""Throughput of a run-time compiled density (logistic regression from CUDA source) on the GPU
and of its host twin under the CPU oracle.  Informational; prints one JSON line per arm."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nutpie_b200
from nutpie_b200 import _lib
from tests import custom_densities as CD

N, D, CH = int(os.environ.get("N_OBS", 2000)), int(os.environ.get("DIM", 48)), int(os.environ.get("CHAINS", 1024))
data = CD.logreg_data(N, D)
gm = nutpie_b200.from_cuda_source(D, CD.LOGREG, data=data, scratch=N)
STAT = {n: i for i, n in enumerate(_lib.STAT_NAMES)}
for rep in range(2):
    s = _lib.PyNutsSettings.Diag(1)
    s.update({"num_tune": 300, "num_draws": 300})
    t = time.time()
    smp = _lib.PySampler(s, gm, n_chains=CH)
    smp.wait()
    wall = time.time() - t
    tr = smp.take_results()
    ms = smp.kernel_ms()
    geo = smp.geometry()
    smp.close()
    steps = tr.stats[..., STAT["n_steps"]].sum()
    print(json.dumps(dict(arm="gpu_custom_logreg", rep=rep, n_obs=N, dim=D, chains=CH, grad_evals=float(steps),
                          kernel_ms=ms, wall_s=wall, evals_per_s_kernel=steps / ms * 1e3,
                          evals_per_s_wall=steps / wall, geometry=geo,
                          mean_steps=float(tr.stats[..., STAT["n_steps"]].mean()),
                          div=float(tr.stats[..., STAT["diverging"]].sum()))))
if os.environ.get("ORACLE", "1") == "1":
    from oracle import pyoracle as O
    from bench import host_cores
    om = O.Model("logreg", D, data=data)
    so = O.default_settings(seed=1, num_tune=300, num_draws=300)
    nthr = host_cores()
    t = time.time()
    ref = O.sample(om, so, 2 * nthr, n_threads=nthr)
    wall = time.time() - t
    print(json.dumps(dict(arm="cpu_oracle_logreg", threads=nthr, chains=2 * nthr, grad_evals=ref["total_steps"],
                          wall_s=wall, evals_per_s=ref["total_steps"] / wall)))
