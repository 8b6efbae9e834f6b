"""adaptation="low_rank" on the GPU vs the CPU oracle: wall time of a whole job (tuning with metric
refreshes + sampling), leapfrogs per draw against the diagonal adaptation.  Informational; one JSON
line per arm.  WORKLOAD = logreg (correlated logistic regression, run-time compiled density) | radon."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nutpie_b200
from nutpie_b200 import _lib
from tests import custom_densities as CD

WORK = os.environ.get("WORKLOAD", "logreg")
CH = int(os.environ.get("CHAINS", 1024))
TUNE, DRAWS = int(os.environ.get("TUNE", 800)), int(os.environ.get("DRAWS", 400))
CUTOFF = float(os.environ.get("CUTOFF", 2.0))
if WORK == "logreg":
    rng = np.random.default_rng(3)
    n, d = 500, 24
    zz = rng.normal(size=(n, d))
    for a in range(0, d - 1, 2):  # correlated pairs of predictors
        zz[:, a + 1] = zz[:, a] * 0.95 + 0.1 * zz[:, a + 1]
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-zz @ rng.normal(size=d)))).astype(float)
    flat = np.concatenate([[n, d], zz.ravel(), y])
    gm = nutpie_b200.from_cuda_source(d, CD.LOGREG, data=flat, scratch=n)
    okw = dict(kind="logreg", dim=d, data=flat)
else:
    dd = nutpie_b200.make_radon_data()
    gm = nutpie_b200.radon_model(dd["y"], dd["county"], dd["floor"], 85)
    okw = dict(kind="radon", dim=175, y=dd["y"], county=dd["county"], floor=dd["floor"], n_county=85)
STAT = {n: i for i, n in enumerate(_lib.STAT_NAMES)}
for adapt in ("diag", "low_rank"):
    s = _lib.PyNutsSettings.LowRank(1) if adapt == "low_rank" else _lib.PyNutsSettings.Diag(1)
    s.update({"num_tune": TUNE, "num_draws": DRAWS})
    if adapt == "low_rank":
        s.mass_matrix_eigval_cutoff = CUTOFF
    t = time.time()
    smp = _lib.PySampler(s, gm, n_chains=CH)
    smp.wait()
    wall = time.time() - t
    tr = smp.take_results()
    ms = smp.kernel_ms()
    smp.close()
    post = tr.stats[:, TUNE:]
    print(json.dumps(dict(arm="gpu", workload=WORK, adaptation=adapt, chains=CH, tune=TUNE, draws=DRAWS,
                          kernel_ms=ms, wall_s=wall, grad_evals=float(tr.stats[..., STAT["n_steps"]].sum()),
                          n_steps_post=float(post[..., STAT["n_steps"]].mean()),
                          step_size_post=float(post[..., STAT["step_size"]].mean()),
                          div_post=float(post[..., STAT["diverging"]].sum()),
                          draws_per_s_post=CH * DRAWS / (ms * 1e-3))), flush=True)
if os.environ.get("ORACLE", "1") == "1":
    from oracle import pyoracle as O
    from bench import host_cores

    kind, dim = okw.pop("kind"), okw.pop("dim")
    om = O.Model(kind, dim, **okw)
    nthr = host_cores()
    nch = nthr if WORK == "radon" else 4 * nthr
    for adapt in (0, 1):
        so = O.default_settings(seed=1, num_tune=TUNE, num_draws=DRAWS, adaptation=adapt,
                                mass_matrix_update_freq=10 if adapt else 1, mass_matrix_eigval_cutoff=CUTOFF)
        t = time.time()
        ref = O.sample(om, so, nch, n_threads=nthr)
        wall = time.time() - t
        post = ref["stats"][:, TUNE:]
        print(json.dumps(dict(arm="cpu_oracle", workload=WORK, adaptation="low_rank" if adapt else "diag",
                              threads=nthr, chains=nch, wall_s=wall, chains_per_s=nch / wall,
                              n_steps_post=float(post[..., 9].mean()), step_size_post=float(post[..., 7].mean()))),
              flush=True)
