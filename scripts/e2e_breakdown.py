"""Where does the end-to-end time of one sampling job go? (create / run / copy / destroy)"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import nutpie_b200
from nutpie_b200 import _lib
d = nutpie_b200.make_radon_data(); model = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], 85)
n, rows, D = 1024, 2000, 175
pd_, ps_ = _lib.PinnedArray((rows, n, D)), _lib.PinnedArray((rows, n, 16))
bufs = {"draws": pd_.array, "stats": ps_.array}
def mk(seed):
    s = _lib.PyNutsSettings.Diag(seed); s.update({"num_tune": 1000, "num_draws": 1000, "init_radius": 1.0}); return s
for stream in (False, True, False, True):
    t0 = time.perf_counter()
    smp = _lib.PySampler(mk(1), model, n_chains=n, autostart=False, trace_buffers=bufs if stream else None)
    t1 = time.perf_counter(); smp.start(); smp.wait(); t2 = time.perf_counter()
    tr = smp.take_results(bufs); t3 = time.perf_counter()
    kms = smp.kernel_ms(); smp.close(); t4 = time.perf_counter()
    print(f"stream={stream}: create {1e3*(t1-t0):.1f} ms, run {1e3*(t2-t1):.1f} ms (kernel {kms:.1f}), copy {1e3*(t3-t2):.1f} ms, destroy {1e3*(t4-t3):.1f} ms, total {1e3*(t4-t0):.1f} ms")
