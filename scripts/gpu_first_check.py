"""First GPU check: CUDA engine vs the CPU oracle on small cases (debug helper)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import pyoracle as O
from nutpie_b200 import _lib, models
from nutpie_b200.datasets import make_radon_data

def settings_pair(**kw):
    s = _lib.PyNutsSettings.Diag(kw.pop("seed", 1))
    so = O.default_settings(seed=s.seed)
    for k, v in kw.items():
        setattr(s._c, k, v); setattr(so, k, v)
    return s, so

def run_gpu(s, model, n_chains, **kw):
    smp = _lib.PySampler(s, model, n_chains=n_chains, **kw)
    smp.wait()
    tr = smp.take_results()
    info = dict(ms=smp.kernel_ms(), geom=smp.geometry(), launches=smp.launch_count())
    smp.close()
    return tr, info

d = make_radon_data(); J = 85
cases = [
    ("normal", models.normal_model(1), O.Model("normal", 1), {}),
    ("normal10", models.normal_model(10, 3.0, 2.0), O.Model("normal", 10, mu=3.0, sigma=2.0), {}),
    ("funnel", models.funnel_model(9), O.Model("funnel", 9), {}),
    ("radon", models.radon_model(d["y"], d["county"], d["floor"], J),
     O.Model("radon", 2*J+5, y=d["y"], county=d["county"], floor=d["floor"], n_county=J), dict(init_radius=1.0)),
]
for tpc in [32, 64, 128]:
    _lib.set_threads_per_chain(tpc)
    for name, gm, om, extra in cases:
        s, so = settings_pair(seed=5, num_tune=200, num_draws=100, store_gradient=1, **extra)
        n = 8
        t = time.time(); tr, info = run_gpu(s, gm, n); tg = time.time() - t
        ref = O.sample(om, so, n)
        dd = np.abs(tr.draws - ref["draws"]).max(axis=(0, 2))
        ds = np.abs(tr.stats - ref["stats"]).max(axis=(0, 1))
        steps_g = tr.stats[..., 9].sum(); steps_o = ref["stats"][..., 9].sum()
        print(f"tpc={tpc} {name}: kernel {info['ms']:.1f} ms wall {tg:.2f}s geom {info['geom']} steps gpu {steps_g:.0f} oracle {steps_o:.0f}")
        print("   draw diff @0,1,5,20,100,299:", [f"{dd[i]:.2e}" for i in (0, 1, 5, 20, 100, 299)])
        print("   n_steps equal frac:", (tr.stats[..., 9] == ref["stats"][..., 9]).mean(), "depth equal frac", (tr.stats[..., 0] == ref["stats"][..., 0]).mean())
# component checks
_lib.set_threads_per_chain(32)
rng = np.random.default_rng(0)
for name, gm, om, extra in cases:
    D = gm.n_dim
    q = rng.normal(size=(16, D)) * 0.5
    lp, g, rc = _lib.logp_grad(gm, q)
    lpo, go, rco = om.logp_grad(q)
    print(name, "logp rel err", np.abs(lp - lpo).max() / np.abs(lpo).max(), "grad err", np.abs(g - go).max() / np.abs(go).max(), rc.sum(), rco.sum())
print("DONE")
