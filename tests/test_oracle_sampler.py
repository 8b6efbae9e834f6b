"""CPU tests: the oracle sampler as a whole against the reference's statistical pins
and determinism contract."""
import numpy as np
import pytest

from oracle import pyoracle as O


def test_config1_normal_posterior():
    """BASELINE config 1 (tests/test_stan.py:16-24 model): mean 0, sd 1."""
    m = O.Model("normal", 1)
    s = O.default_settings(seed=0, num_tune=400, num_draws=1000, init_kind=1)
    r = O.sample(m, s, 4)
    x = r["draws"][:, 400:, 0]
    assert abs(x.mean()) < 0.1
    assert abs(x.std() - 1.0) < 0.06
    st = r["stats"]
    assert O.stat(st, "tuning")[:, :400].all() and not O.stat(st, "tuning")[:, 400:].any()
    assert 0.6 < O.stat(st, "mean_tree_accept")[:, 400:].mean() < 0.95
    assert (O.stat(st, "step_size")[:, 400:] == O.stat(st, "step_size")[:, -1:]).all()  # frozen after tuning


def test_pymc_model_shared_pin():
    """tests/test_pymc.py:397-416: posterior mean of N(-0.1, 1)^3 within 0.05 and of
    N(10, 3)^3 within 0.5 — restated with the analytic device-model twins."""
    s = O.default_settings(seed=1, num_tune=400, num_draws=1000)
    r = O.sample(O.Model("normal", 3, mu=-0.1, sigma=1.0), s, 4)
    np.testing.assert_allclose(r["draws"][:, 400:].mean(), -0.1, atol=0.05)
    r = O.sample(O.Model("normal", 3, mu=10.0, sigma=3.0), s, 4)
    np.testing.assert_allclose(r["draws"][:, 400:].mean(), 10.0, atol=0.5)


def test_seed_contract():
    """tests/test_stan.py:67-101: same seed -> identical, other seed -> different,
    all chains pairwise different."""
    m = O.Model("normal", 2)
    a = O.sample(m, O.default_settings(seed=42, num_tune=50, num_draws=50), 3)
    b = O.sample(m, O.default_settings(seed=42, num_tune=50, num_draws=50), 3, n_threads=1)
    c = O.sample(m, O.default_settings(seed=43, num_tune=50, num_draws=50), 3)
    assert np.array_equal(a["draws"], b["draws"]) and np.array_equal(a["stats"], b["stats"])
    assert not np.array_equal(a["draws"], c["draws"])
    for i in range(3):
        for j in range(i + 1, 3):
            assert not np.allclose(a["draws"][i], a["draws"][j])


def test_chain_offset_reproduces_global_run():
    """chains keyed by GLOBAL id: shards reproduce the unsharded run."""
    m = O.Model("funnel", 5)
    s = O.default_settings(seed=9, num_tune=60, num_draws=40)
    full = O.sample(m, s, 6)
    part = O.sample(m, s, 3, chain_id_offset=3)
    assert np.array_equal(full["draws"][3:], part["draws"])


def test_repeated_values_are_stored():
    """tests/reference/test_deterministic_sampling_numba.txt:3-4 shows repeated draws:
    a transition that selects index 0 repeats the previous point."""
    m = O.Model("normal", 1)
    s = O.default_settings(seed=3, num_tune=100, num_draws=300)
    r = O.sample(m, s, 2)
    x = r["draws"][:, 100:, 0]
    idx = O.stat(r["stats"], "index_in_trajectory")[:, 100:]
    rep = x[:, 1:] == x[:, :-1]
    assert rep.any()
    assert (idx[:, 1:][rep] == 0).all()


def halfnormal_replicate_check(a, gold):
    """`a`: [R, 2, 100] draws of HalfNormal(1) from R independent replicates of the reference's
    golden run (2 chains x (100 tune + 100 draws)); `gold`: the reference's 200 values.
    The reference's seeded stream (rand ChaCha8 + PyMC's jittered init) cannot be reproduced, so
    its file is treated as ONE replicate: each of its summary statistics must be a plausible
    draw from the replicate distribution of OUR sampler (two-sided, 1e-3 per tail), and our
    pooled draws must have the exact half-normal law.  Returns the percentiles for reporting."""
    from scipy import stats as ss

    R = a.shape[0]
    g = gold.reshape(2, 100)

    def ac1(x):
        x = np.log(x)
        x = x - x.mean(-1, keepdims=True)
        return (x[..., 1:] * x[..., :-1]).sum(-1) / (x * x).sum(-1)

    flat = a.reshape(R, 200)
    stat_fns = {
        "mean": lambda v: v.reshape(-1, 200).mean(1),
        "sd": lambda v: v.reshape(-1, 200).std(1),
        "median": lambda v: np.median(v.reshape(-1, 200), axis=1),
        "max": lambda v: v.reshape(-1, 200).max(1),
        "n_repeated": lambda v: (v.reshape(-1, 2, 100)[..., 1:] == v.reshape(-1, 2, 100)[..., :-1]).sum((1, 2)),
        "lag1_autocorr_log": lambda v: ac1(v.reshape(-1, 2, 100)).mean(1),
    }
    pct = {}
    for name, f in stat_fns.items():
        ours, ref = f(a), f(g[None])[0]
        pct[name] = float(((ours < ref).mean() + (ours <= ref).mean()) / 2)
        assert 1e-3 <= pct[name] <= 1 - 1e-3, (name, ref, pct[name], np.percentile(ours, [0.1, 50, 99.9]))
    # exact law of the pooled draws: |N(0, 1)|; thinned to roughly independent draws
    pooled = flat[:, ::10].ravel()
    assert abs(pooled.mean() - np.sqrt(2 / np.pi)) < 4 * 0.603 / np.sqrt(pooled.size)
    assert abs(pooled.std() - np.sqrt(1 - 2 / np.pi)) < 0.02
    assert ss.kstest(pooled[:: max(1, pooled.size // 4000)], "halfnorm").pvalue > 1e-3
    return pct


def test_halfnormal_golden_is_a_plausible_replicate_of_our_sampler():
    """tests/reference/test_deterministic_sampling_numba.txt (tests/test_pymc.py:533-541):
    `pm.HalfNormal("a")`, seed=123, draws=100, tune=100, 2 chains -> 200 values.  Our sampler
    runs the same model (log-transformed half-normal, PyMC's U(-1, 1) jitter) 1500 times with
    the same run shape; see halfnormal_replicate_check for what is asserted.  (For the record:
    the reference's file sits low — mean 0.56 at about the 0.2 percentile of replicate means,
    one chain wandered to a = 3e-4 — but inside the bands.)"""
    from pathlib import Path

    gold = np.loadtxt(Path(__file__).parent / "golden" / "halfnormal_reference_values.txt")
    assert gold.shape == (200,)
    R = 1500
    m = O.Model("halfnormal", 1)
    s = O.default_settings(seed=123, num_tune=100, num_draws=100, init_radius=1.0)
    r = O.sample(m, s, 2 * R)
    a = np.exp(r["draws"][:, 100:, 0]).reshape(R, 2, 100)
    halfnormal_replicate_check(a, gold)


def test_window_schedule_mass_matrix_frozen_late():
    """docs/sample-stats.qmd:85-88: the mass matrix is frozen for the final part of
    tuning (step_size_window = 0.15 of num_tune here)."""
    m = O.Model("normal", 4, mu=0.0, sigma=3.0)
    s = O.default_settings(seed=5, num_tune=200, num_draws=20, store_mass_matrix=1)
    r = O.sample(m, s, 2)
    mm = r["mass_matrix_inv"]
    final = 200 - int(np.ceil(0.15 * 200))
    assert (mm[:, final + 1:] == mm[:, final + 1:final + 2]).all()
    assert not (mm[:, 5] == mm[:, final]).all()
    # adapted diagonal ~ sqrt(Var q / Var g) = sigma^2 for a Gaussian
    np.testing.assert_allclose(mm[:, -1].mean(), 9.0, rtol=0.35)


def test_radon_recovers_truth(radon_data):
    d = radon_data
    J = d["n_county"]; D = 2 * J + 5
    m = O.Model("radon", D, y=d["y"], county=d["county"], floor=d["floor"], n_county=J)
    s = O.default_settings(seed=2, num_tune=300, num_draws=300, init_radius=1.0)
    r = O.sample(m, s, 4)
    dr = r["draws"][:, 300:]
    t = d["truth"]
    assert abs(dr[..., 0].mean() - t["intercept"]) < 0.2
    assert abs(dr[..., J + 2].mean() - t["floor_effect"]) < 0.25
    assert abs(np.exp(dr[..., 2 * J + 4]).mean() - t["sigma"]) < 0.1
    st = r["stats"]
    assert O.stat(st, "diverging")[:, 300:].mean() < 0.02
    # sanity band from docs/_freeze/stan-usage (step ~0.45, ~7 steps/draw on the Stan
    # radon variant): same order of magnitude here
    assert 0.15 < O.stat(st, "step_size")[:, -1].mean() < 1.0
    assert 3 <= O.stat(st, "n_steps")[:, 300:].mean() <= 40
