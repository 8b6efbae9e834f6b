"""Reproduce bench.py's sequence (device arm -> e2e arm) with phase timings of the e2e calls."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import nutpie_b200
from nutpie_b200 import _lib
d = nutpie_b200.make_radon_data(); model = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], 85)
n, rows, D = 1024, 2000, 175
def mk(seed):
    s = _lib.PyNutsSettings.Diag(seed); s.update({"num_tune": 1000, "num_draws": 1000, "num_chains": n, "init_radius": 1.0}); return s
pd_, ps_ = _lib.PinnedArray((rows, n, D)), _lib.PinnedArray((rows, n, 16))
bufs = {"draws": pd_.array, "stats": ps_.array}
variant = sys.argv[1] if len(sys.argv) > 1 else "full"
if variant in ("full", "noess"):
    samplers = [_lib.PySamplerDeferred(mk(100 + i), model, n_chains=n) for i in range(4)]
    for s in samplers: s.start(); s.wait()
    for i, s in enumerate(samplers):
        tr = s.take_results()
        if variant == "full" and i == 3:
            from nutpie_b200.diagnostics import ess
            print("ess", float(ess(tr.draws[:, 1000:, :], max_chains=128).min()))
    for s in samplers: s.close()
    samplers.clear()
T = {}
def timed(name, f):
    def w(*a, **k):
        t = time.perf_counter(); r = f(*a, **k); T[name] = T.get(name, 0) + time.perf_counter() - t; return r
    return w
_lib.PySampler.__init__ = timed("create+start", _lib.PySampler.__init__); _lib.PySampler.wait = timed("wait", _lib.PySampler.wait)
_lib.PySampler.take_results = timed("take", _lib.PySampler.take_results); _lib.PySampler.close = timed("close", _lib.PySampler.close)
for i in range(4):
    T.clear(); t0 = time.perf_counter()
    tr = nutpie_b200.sample(model, draws=1000, tune=1000, chains=n, seed=500 + i, init_radius=1.0, return_raw_trace=True,
                            progress_bar=False, trace_buffers=bufs, expand_on_device=False)
    dt = time.perf_counter() - t0
    print(f"[{variant}] sample() {1e3*dt:.1f} ms :: " + ", ".join(f"{k} {1e3*v:.1f}" for k, v in T.items()), flush=True)
