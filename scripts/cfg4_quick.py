"""Config 4 (iid normal D = 10 000, 512 chains, 200 + 200 draws): kernel time and HBM fraction."""
import sys
sys.path.insert(0, ".")
import json
import nutpie_b200
from nutpie_b200 import _lib
D, C = 10000, 512
if __import__("os").environ.get("STAGE_MODE"):  # nb200_set_stage_loads bit mask (4 = L2 eviction hints on the bulk copies)
    _lib.set_stage_loads(int(__import__("os").environ["STAGE_MODE"]))
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6555.2
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    s = _lib.PyNutsSettings.Diag(7)
    s.update({"num_tune": 200, "num_draws": 200, "store_dims": 16})
    smp = _lib.PySamplerDeferred(s, nutpie_b200.normal_model(D), n_chains=C)
    smp.start(); smp.wait()
    tr = smp.take_results(); ms = smp.kernel_ms(); steps = tr.stats[..., 9].sum(); g = smp.geometry(); smp.close()
    gbs = 72.0 * D * steps / (ms / 1e3) / 1e9
    print(f"config4: {ms:.1f} ms {steps/ms*1e3:.3e} evals/s {gbs:.0f} GB/s algorithmic = {gbs/peak:.3f} of {peak} | {g}", flush=True)
