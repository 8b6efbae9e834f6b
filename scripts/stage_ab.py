"""A/B of the streaming leapfrog modes on config 4 (D = 10000 iid normal, 512 chains,
200 + 200 draws).  Mode bits (nb200_set_stage_loads): 1 bulk-copy staging, 2 alternating sweep
direction, 4 L2 eviction hints.  Modes that share bit 2 must give bit-identical traces."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nutpie_b200
from nutpie_b200 import _lib

D, CH = int(os.environ.get("DIM", 10000)), int(os.environ.get("CHAINS", 512))
MODES = [int(x) for x in os.environ.get("MODES", "0,1,3,5,7").split(",")]
STAT = {n: i for i, n in enumerate(_lib.STAT_NAMES)}
_lib.set_threads_per_chain(int(os.environ.get("TPC", 0)))
gm = nutpie_b200.normal_model(D)
out, best = {}, {}
for rep in range(int(os.environ.get("REPS", 2))):
    for mode in MODES:
        _lib.set_stage_loads(mode)
        s = _lib.PyNutsSettings.Diag(1)
        s.update({"num_tune": 200, "num_draws": 200})
        s._c.store_dims = 16
        smp = _lib.PySampler(s, gm, n_chains=CH)
        smp.wait()
        tr = smp.take_results()
        ms, geo = smp.kernel_ms(), smp.geometry()
        smp.close()
        steps = float(tr.stats[..., STAT["n_steps"]].sum())
        print(json.dumps(dict(mode=mode, rep=rep, dim=D, chains=CH, kernel_ms=ms, evals_per_s=steps / ms * 1e3,
                              algorithmic_GBps=72.0 * D * steps / ms / 1e6, geometry=geo)), flush=True)
        out[mode] = (np.array(tr.draws), np.array(tr.stats))
        best[mode] = max(best.get(mode, 0.0), steps / ms * 1e3)
for a, b in ((0, 1), (1, 5), (3, 7), (1, 3)):
    if a in out and b in out:
        print(f"modes {a} vs {b}: draws identical {np.array_equal(out[a][0], out[b][0])}, "
              f"stats identical {np.array_equal(out[a][1], out[b][1])}, "
              f"max |d draws| {np.abs(out[a][0] - out[b][0]).max():.3e}, "
              f"n_steps sums {out[a][1][..., STAT['n_steps']].sum():.0f} / {out[b][1][..., STAT['n_steps']].sum():.0f}")
for m_, (dr, st_) in out.items():
    post = dr[:, dr.shape[1] // 2:]
    print(f"mode {m_}: n_steps {st_[..., STAT['n_steps']].sum():.0f}  step size {st_[:, -1, STAT['step_size']].mean():.4f}  "
          f"draw mean {post.mean():+.4f} var {post.var():.4f}  divergences {st_[..., STAT['diverging']].sum():.0f}  "
          f"lib {os.environ.get('NB200_LIB', 'default')} tpc {os.environ.get('TPC', 'auto')}")
winner = max(best, key=best.get)
print("best mode", winner, best)
if os.environ.get("BEST_FILE"):
    open(os.environ["BEST_FILE"], "w").write(str(winner))
