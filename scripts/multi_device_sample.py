"""One process, every GPU of the box: nutpie_b200.sample(model, chains=1024 x n_gpus, devices="all")
(PyMultiSampler: a sampler, stream and waiting host thread per device) — wall time per call."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import nutpie_b200
from nutpie_b200 import _lib
n = _lib.device_count()
d = nutpie_b200.make_radon_data(); model = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], 85)
for devices, chains in ((1, 1024), ("all", 1024 * n)):
    for i in range(3):
        t0 = time.perf_counter()
        res = nutpie_b200.sample(model, draws=1000, tune=1000, chains=chains, seed=900 + i, init_radius=1.0,
                                 progress_bar=False, devices=devices)
        dt = time.perf_counter() - t0
        steps = res.sample_stats["n_steps"].sum() + res.warmup_sample_stats["n_steps"].sum()
        print(f"devices={devices} ({n if devices == 'all' else 1} GPU): {chains} chains, call {i}: {1e3*dt:.0f} ms, "
              f"{steps/dt:.3e} grad evals/s, posterior county_effect {res.posterior['county_effect'].shape}", flush=True)
        del res
