"""The DEFAULT call — nutpie_b200.sample(model, chains=1024, tune=1000, draws=1000) with no caller
buffers: wall time per call (first = cold: staging ring, page faults on fresh arrays)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import nutpie_b200
d = nutpie_b200.make_radon_data(); model = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], 85)
for i in range(4):
    t0 = time.perf_counter()
    res = nutpie_b200.sample(model, draws=1000, tune=1000, chains=1024, seed=900 + i, init_radius=1.0, progress_bar=False)
    dt = time.perf_counter() - t0
    n = res.sample_stats["n_steps"].sum() + res.warmup_sample_stats["n_steps"].sum()
    x = res.posterior["county_effect"]
    print(f"default sample() call {i}: {1e3*dt:.1f} ms, {n/dt:.3e} grad evals/s, county_effect {x.shape} mean {x.mean():+.4f}", flush=True)
    del res, x
