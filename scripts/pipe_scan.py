"""Kernel time of the radon sampler against chains per SM (148 x k chains), one-warp vs two-warp
kernel: shows where per-SM resources (issue slots, L1, registers) start to couple the chains."""
import sys
sys.path.insert(0, ".")
import numpy as np
import nutpie_b200
from nutpie_b200 import _lib

d = nutpie_b200.make_radon_data()
m = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], d["n_county"])
tune = draws = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for k in (1, 2, 3, 4, 5, 6, 7):
    out = []
    for pipe in (False, True):
        _lib.set_pipeline(pipe)
        s = _lib.PyNutsSettings.Diag(77)
        s.update({"num_tune": tune, "num_draws": draws, "init_radius": 1.0})
        smp = _lib.PySampler(s, m, n_chains=148 * k)
        smp.wait()
        tr = smp.take_results()
        out.append((smp.kernel_ms(), tr.stats[..., 9].sum(), smp.geometry()))
        smp.close()
    (a, sa, ga), (b, sb, gb) = out
    print(f"k={k}: one-warp {a:7.1f} ms ({sa/a*1e3:.3e}/s, smem_slots {ga['smem_slots']}) | piped {b:7.1f} ms "
          f"({sb/b*1e3:.3e}/s, smem_slots {gb['smem_slots']}, block {gb['block']}) | x{a/b:.3f}", flush=True)
