"""A/B of the streaming leapfrog with and without bulk-copy staging on config 4
(D = 10000 iid normal, 512 chains, 200 + 200 draws); results must be bit-identical."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nutpie_b200
from nutpie_b200 import _lib

D, CH = int(os.environ.get("DIM", 10000)), int(os.environ.get("CHAINS", 512))
STAT = {n: i for i, n in enumerate(_lib.STAT_NAMES)}
gm = nutpie_b200.normal_model(D)
out = {}
for rep in range(int(os.environ.get("REPS", 2))):
    for stage in (0, 1):
        _lib.set_stage_loads(bool(stage))
        s = _lib.PyNutsSettings.Diag(1)
        s.update({"num_tune": 200, "num_draws": 200})
        s._c.store_dims = 16
        smp = _lib.PySampler(s, gm, n_chains=CH)
        smp.wait()
        tr = smp.take_results()
        ms, geo = smp.kernel_ms(), smp.geometry()
        smp.close()
        steps = float(tr.stats[..., STAT["n_steps"]].sum())
        print(json.dumps(dict(stage=stage, rep=rep, dim=D, chains=CH, kernel_ms=ms, evals_per_s=steps / ms * 1e3,
                              algorithmic_GBps=72.0 * D * steps / ms / 1e6, geometry=geo)), flush=True)
        out[stage] = (np.array(tr.draws), np.array(tr.stats))
print("bit-identical draws:", np.array_equal(out[0][0], out[1][0]),
      "stats:", np.array_equal(out[0][1], out[1][1]))
