"""Sub-warp geometry (4/8/16 lanes per chain) against a warp per chain on the small-D configs:
BASELINE config 5 (funnel D = 9, 4096 chains, maxdepth 12) and config 1 (normal D = 1)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import nutpie_b200
from nutpie_b200 import _lib

def run(model, chains, tpc, upd, seed=11):
    _lib.set_threads_per_chain(tpc)
    s = _lib.PyNutsSettings.Diag(seed)
    s.update(upd)
    smp = _lib.PySampler(s, model, n_chains=chains)
    smp.wait()
    tr = smp.take_results()
    ms, g = smp.kernel_ms(), smp.geometry()
    smp.close()
    return tr, ms, g

for name, model, chains, upd in (
        ("funnel D=9 4096 chains maxdepth 12", nutpie_b200.funnel_model(9), 4096, {"num_tune": 1000, "num_draws": 1000, "maxdepth": 12}),
        ("normal D=1 4096 chains", nutpie_b200.normal_model(1), 4096, {"num_tune": 400, "num_draws": 1000}),
        ("normal D=1 4 chains (config 1)", nutpie_b200.normal_model(1), 4, {"num_tune": 400, "num_draws": 1000})):
    base = None
    for tpc in (32, 16, 8, 4, 0):
        try:
            tr, ms, g = run(model, chains, tpc, upd)
        except Exception as exc:
            print(f"{name}: tpc {tpc}: {exc}")
            continue
        st = tr.stats
        steps = st[..., 9].sum()
        line = (f"{name}: tpc {tpc:2d} -> lanes {g['threads_per_chain']:3d} block {g['block']} grid {g['grid']} smem_slots {g['smem_slots']}: "
                f"{ms:8.1f} ms {steps/ms*1e3:.3e}/s  mean n_steps {st[:, -500:, 9].mean():.2f} div {st[:, -500:, 6].mean():.4f} "
                f"step {st[:, -1, 7].mean():.4f} mean x0 {tr.draws[:, -500:, 0].mean():+.4f} sd {tr.draws[:, -500:, 0].std():.4f}")
        if base is None:
            base = tr
        else:
            line += f" | n_steps equal to tpc32: {np.mean(tr.stats[..., 9] == base.stats[..., 9]):.4f}"
        print(line, flush=True)
