"""Time nutpie_b200.sample() (the bench's e2e call) and its phases."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import nutpie_b200
from nutpie_b200 import _lib
d = nutpie_b200.make_radon_data(); model = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], 85)
n, rows, D = 1024, 2000, 175
pd_, ps_ = _lib.PinnedArray((rows, n, D)), _lib.PinnedArray((rows, n, 16))
bufs = {"draws": pd_.array, "stats": ps_.array}
orig_init = _lib.PySampler.__init__; orig_wait = _lib.PySampler.wait; orig_take = _lib.PySampler.take_results; orig_close = _lib.PySampler.close
T = {}
def timed(name, f):
    def w(*a, **k):
        t = time.perf_counter(); r = f(*a, **k); T[name] = T.get(name, 0) + time.perf_counter() - t; return r
    return w
_lib.PySampler.__init__ = timed("create+start", orig_init); _lib.PySampler.wait = timed("wait", orig_wait)
_lib.PySampler.take_results = timed("take", orig_take); _lib.PySampler.close = timed("close", orig_close)
for i in range(4):
    T.clear(); t0 = time.perf_counter()
    tr = nutpie_b200.sample(model, draws=1000, tune=1000, chains=n, seed=500 + i, init_radius=1.0, return_raw_trace=True,
                            progress_bar=False, trace_buffers=bufs)
    dt = time.perf_counter() - t0
    print(f"sample() {1e3*dt:.1f} ms :: " + ", ".join(f"{k} {1e3*v:.1f}" for k, v in T.items()), flush=True)
    if i == 1:
        import torch; torch.zeros(1, device="cuda"); print("torch cuda initialised")
