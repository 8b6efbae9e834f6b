"""Short run of one workload for ncu (argv[1] = radon|cfg4|funnel)."""
import sys
sys.path.insert(0, ".")
import nutpie_b200
from nutpie_b200 import _lib
which = sys.argv[1] if len(sys.argv) > 1 else "radon"
tpc = int(sys.argv[2]) if len(sys.argv) > 2 else 0
_lib.set_threads_per_chain(tpc)
import os
if os.environ.get("STAGE_MODE"):
    _lib.set_stage_loads(int(os.environ["STAGE_MODE"]))
if which == "radon":
    d = nutpie_b200.make_radon_data()
    model = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], 85); n = 1024; upd = {"num_tune": 30, "num_draws": 10, "init_radius": 1.0}
elif which == "cfg4":
    model = nutpie_b200.normal_model(10000); n = 512; upd = {"num_tune": int(os.environ.get("PROF_TUNE", 40)), "num_draws": 10, "store_dims": 16}
else:
    model = nutpie_b200.funnel_model(9); n = 4096; upd = {"num_tune": 30, "num_draws": 10, "maxdepth": 12}
s = _lib.PyNutsSettings.Diag(3); s.update(upd)
smp = _lib.PySamplerDeferred(s, model, n_chains=n); smp.start(); smp.wait()
tr = smp.take_results(); print(which, "steps", tr.stats[..., 9].sum(), "ms", smp.kernel_ms(), smp.geometry()); smp.close()
