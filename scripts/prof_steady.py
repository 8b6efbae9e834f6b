"""Radon run cut into launches of 500 draws (tune 1000 + draws 1000 -> 4 launches): the last
launch is pure steady-state sampling.  Profile it with `ncu --launch-skip 3 --launch-count 1`.
argv[1] = 0/1 pipeline off/on."""
import sys
sys.path.insert(0, ".")
import nutpie_b200
from nutpie_b200 import _lib
pipe = int(sys.argv[1]) if len(sys.argv) > 1 else 1
_lib.set_pipeline(bool(pipe))
d = nutpie_b200.make_radon_data()
model = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], 85)
s = _lib.PyNutsSettings.Diag(3)
s.update({"num_tune": 1000, "num_draws": 1000, "init_radius": 1.0})
smp = _lib.PySamplerDeferred(s, model, n_chains=1024, draws_per_launch=500)
smp.start(); smp.wait()
tr = smp.take_results()
st = tr.stats
print("pipe", pipe, "steps", st[..., 9].sum(), "ms", smp.kernel_ms(), "launches", smp.launch_count(), smp.geometry())
for a, b in ((0, 500), (500, 1000), (1000, 1500), (1500, 2000)):
    print(f"  draws {a}-{b}: steps {st[:, a:b, 9].sum():.0f} mean depth {st[:, a:b, 0].mean():.2f} max-chain steps {st[:, a:b, 9].sum(1).max():.0f}")
smp.close()
