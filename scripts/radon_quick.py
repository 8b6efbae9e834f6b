"""Radon, 1024 chains, TUNE+DRAWS draws: kernel-only gradient evaluations/s (A/B of libraries via NB200_LIB)."""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nutpie_b200
from nutpie_b200 import _lib

tune, draws = int(os.environ.get("TUNE", 400)), int(os.environ.get("DRAWS", 400))
d = nutpie_b200.make_radon_data()
model = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], 85)
for rep in range(int(os.environ.get("REPS", 3))):
    s = _lib.PyNutsSettings.Diag(11 + rep)
    s.update({"num_tune": tune, "num_draws": draws, "init_radius": 1.0})
    smp = _lib.PySamplerDeferred(s, model, n_chains=1024)
    smp.start()
    smp.wait()
    tr = smp.take_results()
    ms = smp.kernel_ms()
    steps = tr.stats[..., 9].sum()
    print(f"radon lib={os.environ.get('NB200_LIB', 'default')} rep {rep}: {steps / ms * 1e3:.4g} evals/s  {ms:.1f} ms  "
          f"steps {steps:.0f}  div {tr.stats[..., 6].sum():.0f}  "
          f"sha1(draws) {hashlib.sha1(np.ascontiguousarray(tr.draws).tobytes()).hexdigest()[:12]}  "
          f"sha1(stats) {hashlib.sha1(np.ascontiguousarray(tr.stats).tobytes()).hexdigest()[:12]}", flush=True)
    smp.close()
