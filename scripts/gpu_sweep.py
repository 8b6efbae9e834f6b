"""Geometry sweep on the benchmark workload (short runs): threads per chain, smem slots,
chains per block.  Prints grad evals/s for each."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import nutpie_b200
from nutpie_b200 import _lib

d = nutpie_b200.make_radon_data(); J = 85
model = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], J)

def run(n_chains, tune, draws, tpc, slots, cpb=0, model=model, **upd):
    _lib.set_threads_per_chain(tpc); _lib.set_smem_slots(slots); _lib.set_chains_per_block(cpb)
    s = _lib.PyNutsSettings.Diag(3)
    s.update({"num_tune": tune, "num_draws": draws, "init_radius": 1.0, **upd})
    smp = _lib.PySamplerDeferred(s, model, n_chains=n_chains)
    smp.start(); smp.wait()
    tr = smp.take_results(); ms = smp.kernel_ms(); g = smp.geometry(); smp.close()
    steps = tr.stats[..., 9].sum()
    return steps / ms * 1e3, ms, g

which = sys.argv[1] if len(sys.argv) > 1 else "radon"
if which == "radon":
    run(64, 50, 50, 32, -1)
    for tpc, slots, cpb, unroll in [(32, -1, 0, 1), (32, -1, 0, 0), (32, 8, 0, 1), (32, 2, 0, 1), (64, -1, 0, 1), (64, 8, 0, 1),
                                    (128, -1, 0, 1), (128, 8, 0, 1), (128, -1, 0, 0), (256, -1, 0, 1)]:
        _lib.set_unroll(unroll)
        for _ in (0,):
            for _ in (0,):
                try:
                    v, ms, g = run(1024, 300, 200, tpc, slots, cpb)
                    print(f"radon 1024ch tpc={tpc} slots={slots} cpb={cpb} unroll={unroll}: {v:.3e} evals/s  {ms:.1f} ms  {g}", flush=True)
                except Exception as e:
                    print(f"radon tpc={tpc} slots={slots} cpb={cpb}: FAILED {e}", flush=True)
elif which == "cfg4":
    m4 = nutpie_b200.normal_model(10000)
    for tpc in (128, 256):
        v, ms, g = run(512, 200, 200, tpc, -1, model=m4, store_dims=16)
        print(f"cfg4 tpc={tpc}: {v:.3e} evals/s -> {v*72*10000/1e9:.0f} GB/s algorithmic  {ms:.1f} ms {g}", flush=True)
elif which == "funnel":
    mf = nutpie_b200.funnel_model(9)
    for tpc in (32,):
        for cpb in (1, 4, 8):
            v, ms, g = run(4096, 300, 200, tpc, -1, cpb, model=mf, maxdepth=12)
            print(f"funnel 4096ch tpc={tpc} cpb={cpb}: {v:.3e} evals/s {ms:.1f} ms {g}", flush=True)
