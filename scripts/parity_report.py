"""Parity report at BASELINE.json sizes: CUDA engine vs CPU oracle, same settings.
Chains are independent replicates, so means are compared in units of the combined MCSE."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import nutpie_b200
from nutpie_b200 import _lib
from oracle import pyoracle as O

def mcse_z(a, b):
    ma, mb = a.mean(1), b.mean(1)
    se = np.sqrt(ma.var(0, ddof=1) / len(ma) + mb.var(0, ddof=1) / len(mb))
    return np.abs(ma.mean(0) - mb.mean(0)) / se

def report(name, gm, om, n_gpu, n_cpu, tune, draws, low_rank=False, **kw):
    s = _lib.PyNutsSettings.LowRank(11) if low_rank else _lib.PyNutsSettings.Diag(11)
    so = O.default_settings(seed=11)
    if low_rank:
        kw = dict(adaptation=1, mass_matrix_update_freq=10, **kw)
    for k, v in dict(num_tune=tune, num_draws=draws, **kw).items():
        setattr(s._c, k, v); setattr(so, k, v)
    smp = _lib.PySampler(s, gm, n_chains=n_gpu); smp.wait(); tr = smp.take_results(); ms = smp.kernel_ms(); smp.close()
    t = time.time(); ref = O.sample(om, so, n_cpu); tc = time.time() - t
    g, c = tr.draws[:, tune:], ref["draws"][:, tune:]
    z = mcse_z(g, c); sdr = g.std((0, 1)) / c.std((0, 1))
    st, rs = tr.stats[:, tune:], ref["stats"][:, tune:]
    out = {"config": name, "gpu_chains": n_gpu, "cpu_chains": n_cpu, "tune": tune, "draws": draws,
           "max_mean_diff_in_MCSE": float(z.max()), "sd_ratio_min_max": [float(sdr.min()), float(sdr.max())],
           "step_size_gpu_cpu": [float(tr.stats[:, -1, 7].mean()), float(ref["stats"][:, -1, 7].mean())],
           "mean_n_steps_gpu_cpu": [float(st[..., 9].mean()), float(rs[..., 9].mean())],
           "mean_tree_accept_gpu_cpu": [float(st[..., 10].mean()), float(rs[..., 10].mean())],
           "divergence_rate_gpu_cpu": [float(st[..., 6].mean()), float(rs[..., 6].mean())],
           "depth_hist_gpu": (np.bincount(st[..., 0].astype(int).ravel(), minlength=13) / st[..., 0].size).round(4).tolist(),
           "depth_hist_cpu": (np.bincount(rs[..., 0].astype(int).ravel(), minlength=13) / rs[..., 0].size).round(4).tolist(),
           "gpu_kernel_ms": ms, "gpu_grad_evals_per_s": float(tr.stats[..., 9].sum() / ms * 1e3),
           "cpu_seconds": tc, "cpu_grad_evals_per_s": ref["total_steps"] / tc}
    print(json.dumps(out), flush=True)

d = nutpie_b200.make_radon_data(); J = 85
report("1: normal(mu,1), D=1, 4 chains (x64 replicates for MCSE)", nutpie_b200.normal_model(1), O.Model("normal", 1), 256, 256, 400, 1000, init_kind=1)
report("2: radon D=175, 1024 chains", nutpie_b200.radon_model(d["y"], d["county"], d["floor"], J),
       O.Model("radon", 2*J+5, y=d["y"], county=d["county"], floor=d["floor"], n_county=J), 1024, 256, 1000, 1000, init_radius=1.0)
report("4: iid normal D=10000, 512 chains (first 16 coords stored)", nutpie_b200.normal_model(10000), O.Model("normal", 10000), 512, 32, 200, 200, store_dims=16)
report("5: funnel D=9, 4096 chains, maxdepth 12", nutpie_b200.funnel_model(9), O.Model("funnel", 9), 4096, 1024, 1000, 1000, maxdepth=12)
# adaptation="low_rank" (SURVEY §8 f4): correlated logistic regression as run-time compiled CUDA source, and radon
from tests import custom_densities as CD
rng = np.random.default_rng(3)
n_obs, dim = 500, 24
zz = rng.normal(size=(n_obs, dim))
for a in range(0, dim - 1, 2):
    zz[:, a + 1] = zz[:, a] * 0.95 + 0.1 * zz[:, a + 1]
yy = (rng.uniform(size=n_obs) < 1 / (1 + np.exp(-zz @ rng.normal(size=dim)))).astype(float)
flat = np.concatenate([[n_obs, dim], zz.ravel(), yy])
report("low_rank: logistic regression D=24 (NVRTC density), 1024 chains", nutpie_b200.from_cuda_source(dim, CD.LOGREG, data=flat, scratch=n_obs),
       O.Model("logreg", dim, data=flat), 1024, 64, 800, 400, low_rank=True)
report("low_rank: radon D=175, 256 chains, cutoff 4", nutpie_b200.radon_model(d["y"], d["county"], d["floor"], J),
       O.Model("radon", 2*J+5, y=d["y"], county=d["county"], floor=d["floor"], n_county=J), 256, 16, 400, 200, low_rank=True,
       init_radius=1.0, mass_matrix_eigval_cutoff=4.0)
