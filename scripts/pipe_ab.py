"""A/B of the two-warp (integrator + tree) radon kernel against the one-warp kernel:
bit-identical traces, kernel time.  Usage: python scripts/pipe_ab.py [chains] [tune] [draws]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import nutpie_b200
from nutpie_b200 import _lib

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
tune = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
draws = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
d = nutpie_b200.make_radon_data()
m = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], d["n_county"])

def run(pipe, seed):
    _lib.set_pipeline(pipe)
    s = _lib.PyNutsSettings.Diag(seed)
    s.update({"num_tune": tune, "num_draws": draws, "init_radius": 1.0})
    smp = _lib.PySampler(s, m, n_chains=chains)
    smp.wait()
    tr = smp.take_results()
    ms, geom = smp.kernel_ms(), smp.geometry()
    smp.close()
    return tr, ms, geom

for seed in range(100, 100 + reps):
    a, ms_a, ga = run(False, seed)
    b, ms_b, gb = run(True, seed)
    steps = a.stats[..., 9].sum()
    same = np.array_equal(a.draws, b.draws) and np.array_equal(a.stats, b.stats)
    print(f"seed {seed}: one-warp {ms_a:8.1f} ms {steps/ms_a*1e3:.3e}/s | piped {ms_b:8.1f} ms "
          f"{steps/ms_b*1e3:.3e}/s | x{ms_a/ms_b:.3f} | identical={same} | {gb}", flush=True)
