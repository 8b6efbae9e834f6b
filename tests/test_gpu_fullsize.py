"""Parity at BASELINE.json's FULL sizes (SURVEY.md §8d tolerances), run by the driver's
`-m gpu` pass: the CUDA engine at the chain counts / draw counts / tree depths of configs 2-5
against the CPU oracle on the same settings (the oracle runs a subset of the SAME global chain
ids where it is slow), plus the size-independent properties the domain offers: chain sharding
equals the unsharded run bit for bit, elementwise densities reproduce the oracle's trees
exactly.  Tolerances (stated in SURVEY.md §8d): posterior means within 4 MCSE, sds within 5 %,
step size and n_steps within 10 %, depth histogram within 0.03 per bin, divergence rate."""
from pathlib import Path

import numpy as np
import pytest

import nutpie_b200
from nutpie_b200 import _lib
from oracle import pyoracle as O
from tests import custom_densities as CD

pytestmark = pytest.mark.gpu
STAT = {n: i for i, n in enumerate(_lib.STAT_NAMES)}


def settings_pair(seed, **kw):
    s = _lib.PyNutsSettings.Diag(seed)
    so = O.default_settings(seed=seed)
    for k, v in kw.items():
        setattr(s._c, k, v)
        setattr(so, k, v)
    return s, so


def run_gpu(s, model, n_chains, **kw):
    smp = _lib.PySampler(s, model, n_chains=n_chains, **kw)
    try:
        smp.wait()
        tr = smp.take_results()
        return tr, smp.kernel_ms()
    finally:
        smp.close()


def mcse_z(a, b):
    """|mean_a - mean_b| per parameter in units of the combined Monte-Carlo standard error;
    chains are independent replicates, so the MCSE comes from the spread of chain means."""
    ma, mb = a.mean(1), b.mean(1)
    se = np.sqrt(ma.var(0, ddof=1) / len(ma) + mb.var(0, ddof=1) / len(mb))
    return np.abs(ma.mean(0) - mb.mean(0)) / se


def depth_hist(stats, lo, n=14):
    d = stats[:, lo:, STAT["depth"]].astype(int).ravel()
    return np.bincount(d, minlength=n) / d.size


def test_config2_radon_1024_chains_1000_1000(radon_data):
    """BASELINE config 2 at full size; the oracle runs chains 0..255 of the same job."""
    d = radon_data
    J = d["n_county"]
    gm = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], J)
    om = O.Model("radon", 2 * J + 5, y=d["y"], county=d["county"], floor=d["floor"], n_county=J)
    tune = draws = 1000
    s, so = settings_pair(11, num_tune=tune, num_draws=draws, init_radius=1.0)
    tr, ms = run_gpu(s, gm, 1024)
    ref = O.sample(om, so, 256)
    # same chain ids, same streams: the first draws agree to rounding
    dd = np.abs(tr.draws[:256, :3] - ref["draws"][:, :3]).max()
    assert dd < 1e-6, dd
    g, c = tr.draws[:, tune:], ref["draws"][:, tune:]
    z = mcse_z(g, c)
    assert z.max() < 4.0, z.max()
    sdr = g.std((0, 1)) / c.std((0, 1))
    assert np.abs(sdr - 1).max() < 0.05, (sdr.min(), sdr.max())
    st, rs = tr.stats, ref["stats"]
    for name, tol in (("step_size", 0.10), ("n_steps", 0.10), ("mean_tree_accept", 0.03)):
        a, b = st[:, tune:, STAT[name]].mean(), rs[:, tune:, STAT[name]].mean()
        assert abs(a / b - 1) < tol, (name, a, b)
    assert np.abs(depth_hist(st, tune) - depth_hist(rs, tune)).max() < 0.03
    assert st[:, tune:, STAT["diverging"]].mean() < 0.005
    # warm-up as a whole does the same amount of work (adaptation schedule parity)
    a, b = st[:, :tune, STAT["n_steps"]].mean(), rs[:, :tune, STAT["n_steps"]].mean()
    assert abs(a / b - 1) < 0.10, (a, b)
    assert st[..., STAT["n_steps"]].sum() / ms * 1e3 > 5e7  # it is the fast path that ran


def test_config3_eight_shards_equal_one_8192_chain_job(radon_data):
    """BASELINE config 3 on one GPU: eight 1024-chain shards addressed by chain_id_offset (what
    the eight ranks run) — shard k reproduces chains [1024k, 1024k+1024) of the unsharded job
    bit for bit, and the oracle's chains with the same global ids to rounding."""
    d = radon_data
    J = d["n_county"]
    gm = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], J)
    om = O.Model("radon", 2 * J + 5, y=d["y"], county=d["county"], floor=d["floor"], n_county=J)
    tune, draws = 150, 100
    mk = lambda: settings_pair(23, num_tune=tune, num_draws=draws, init_radius=1.0)
    full, _ = run_gpu(mk()[0], gm, 8192)
    total = 0
    for k in (0, 3, 7):
        shard, _ = run_gpu(mk()[0], gm, 1024, chain_id_offset=1024 * k)
        sl = slice(1024 * k, 1024 * (k + 1))
        assert np.array_equal(shard.draws, full.draws[sl]) and np.array_equal(shard.stats, full.stats[sl])
        assert shard.stats[0, 0, STAT["chain"]] == 1024 * k
        ref = O.sample(om, mk()[1], 8, chain_id_offset=1024 * k)
        assert np.abs(shard.draws[:8, :3] - ref["draws"][:, :3]).max() < 1e-6
        np.testing.assert_array_equal(shard.stats[:8, :20, STAT["n_steps"]], ref["stats"][:, :20, STAT["n_steps"]])
        total += shard.stats[..., STAT["n_steps"]].sum()
    assert total > 0
    # every chain of the job is distinct (determinism contract, tests/test_stan.py:67-101)
    last = full.draws[:, -1, 0]
    assert len(np.unique(last)) == 8192


def test_config4_iid_normal_D10000_512_chains():
    """BASELINE config 4 at full size.  The density is elementwise, so the oracle's trees are
    reproduced EXACTLY (depth, n_steps, index in trajectory, divergences) for the chains it
    runs; all 512 chains are checked through properties of the exact posterior."""
    D, C, tune, draws = 10000, 512, 200, 200
    gm, om = nutpie_b200.normal_model(D), O.Model("normal", D)
    s, so = settings_pair(7, num_tune=tune, num_draws=draws, store_dims=16)
    tr, ms = run_gpu(s, gm, C)
    ref = O.sample(om, so, 12)
    for name in ("depth", "n_steps", "index_in_trajectory", "diverging"):
        np.testing.assert_array_equal(tr.stats[:12, :, STAT[name]], ref["stats"][..., STAT[name]], err_msg=name)
    # positions: equal to rounding at first; 10 000-term reductions are summed in a different
    # order on the device, and the adapted step size carries those roundings forward
    np.testing.assert_allclose(tr.draws[:12, :10], ref["draws"][:, :10], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(tr.draws[:12], ref["draws"], rtol=1e-2, atol=1e-4)
    np.testing.assert_allclose(tr.stats[:12, :, STAT["step_size"]], ref["stats"][..., STAT["step_size"]], rtol=1e-5)
    post = tr.draws[:, tune:]                      # 16 stored coordinates of N(0, 1)^D
    assert np.abs(post.mean((0, 1))).max() < 4 / np.sqrt(C * draws / 4)
    assert np.abs(post.std((0, 1)) - 1).max() < 0.03
    # -2 logp ~ chi2(D): mean D, sd sqrt(2 D)
    lp = -2 * tr.stats[:, tune:, STAT["logp"]]
    assert abs(lp.mean() / D - 1) < 0.01 and abs(lp.std() / np.sqrt(2 * D) - 1) < 0.1
    assert tr.stats[:, tune:, STAT["diverging"]].sum() == 0
    acc = tr.stats[:, tune:, STAT["mean_tree_accept"]].mean()
    assert abs(acc - 0.8) < 0.05, acc
    assert tr.stats[..., STAT["n_steps"]].sum() / ms * 1e3 > 2e6


def test_config5_funnel_4096_chains_maxdepth_12():
    """BASELINE config 5 at full size: divergence rate and tree-depth histogram vs the oracle
    (1024 chains of the same job), deep trees present."""
    gm, om = nutpie_b200.funnel_model(9), O.Model("funnel", 9)
    tune = draws = 1000
    s, so = settings_pair(11, num_tune=tune, num_draws=draws, maxdepth=12)
    tr, _ = run_gpu(s, gm, 4096)
    ref = O.sample(om, so, 1024)
    st, rs = tr.stats, ref["stats"]
    gd, od = st[:, tune:, STAT["diverging"]].mean(), rs[:, tune:, STAT["diverging"]].mean()
    assert abs(gd - od) < 0.003 + 0.15 * od, (gd, od)
    hg, ho = depth_hist(st, tune), depth_hist(rs, tune)
    assert np.abs(hg - ho).max() < 0.03, (hg, ho)
    for name, tol in (("n_steps", 0.10), ("step_size", 0.10)):
        a, b = st[:, tune:, STAT[name]].mean(), rs[:, tune:, STAT[name]].mean()
        assert abs(a / b - 1) < tol, (name, a, b)
    assert st[..., STAT["depth"]].max() >= 8      # the stiff neck is explored with deep trees
    z = mcse_z(tr.draws[:, tune:], ref["draws"][:, tune:])
    assert z.max() < 4.5, z


def test_config1_normal_mu_1_four_chains():
    """BASELINE config 1 (the reference's CPU-runnable case): Stan `normal(mu, 1)`, D = 1,
    4 chains, 1000 draws, N(0, 1) initial points, seeds {0, 1, 2} — draw for draw against
    the oracle (the density is order-independent) and against the exact posterior."""
    gm, om = nutpie_b200.normal_model(1), O.Model("normal", 1)
    pooled = []
    for seed in (0, 1, 2):
        s, so = settings_pair(seed, num_tune=400, num_draws=1000, init_kind=1)
        tr, _ = run_gpu(s, gm, 4)
        ref = O.sample(om, so, 4)
        np.testing.assert_array_equal(tr.stats[..., STAT["n_steps"]], ref["stats"][..., STAT["n_steps"]])
        # (the device contracts a*b+c into FMAs, the oracle is built with -ffp-contract=off: the
        # last-bit differences are carried forward by 1400 draws of adaptation)
        np.testing.assert_allclose(tr.draws[:, :50], ref["draws"][:, :50], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(tr.draws, ref["draws"], rtol=1e-3, atol=1e-5)
        pooled.append(tr.draws[:, 400:, 0])
    x = np.concatenate(pooled).ravel()
    assert abs(x.mean()) < 4 / np.sqrt(x.size / 3) and abs(x.std() - 1) < 0.03


def test_halfnormal_golden_replicates_on_the_gpu():
    """The reference's seeded golden file (tests/reference/test_deterministic_sampling_numba.txt)
    as one replicate among 1500 GPU replicates of the same run shape — the device twin of
    tests/test_oracle_sampler.py::test_halfnormal_golden_is_a_plausible_replicate_of_our_sampler."""
    from tests.test_oracle_sampler import halfnormal_replicate_check

    gold = np.loadtxt(Path(__file__).parent / "golden" / "halfnormal_reference_values.txt")
    R = 1500
    gm = nutpie_b200.from_cuda_source(1, CD.HALFNORMAL)
    s, _ = settings_pair(123, num_tune=100, num_draws=100, init_radius=1.0)
    tr, _ = run_gpu(s, gm, 2 * R)
    a = np.exp(tr.draws[:, 100:, 0]).reshape(R, 2, 100)
    halfnormal_replicate_check(a, gold)
    # and chain for chain the oracle's run of the same chains
    om = O.Model("halfnormal", 1)
    ref = O.sample(om, O.default_settings(seed=123, num_tune=100, num_draws=100, init_radius=1.0), 64)
    # (exp() differs in the last bit between libm and the device: a handful of trees flip)
    same = tr.stats[:64, :, STAT["n_steps"]] == ref["stats"][..., STAT["n_steps"]]
    assert same.mean() > 0.99, same.mean()
    np.testing.assert_allclose(tr.draws[:64, :5], ref["draws"][:, :5], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("tpc", [128, 256])
def test_pause_resume_with_a_cta_per_chain_is_bit_identical(tpc):
    """Several warps per chain read the host's stop flag ONCE per draw (thread 0, published
    through shared memory): pausing at arbitrary moments must neither hang nor change a bit."""
    import time

    _lib.set_threads_per_chain(tpc)
    try:
        gm = nutpie_b200.normal_model(4096)
        mk = lambda: settings_pair(5, num_tune=120, num_draws=80, store_dims=8)[0]
        ref, _ = run_gpu(mk(), gm, 64)
        smp = _lib.PySampler(mk(), gm, n_chains=64)
        try:
            n_pauses = 0
            while not smp.is_finished() and n_pauses < 200:
                time.sleep(0.002)
                smp.pause()
                n_pauses += 1
                smp.resume()
            smp.wait()
            tr = smp.take_results()
        finally:
            smp.close()
        assert n_pauses >= 1
        assert np.array_equal(tr.draws, ref.draws) and np.array_equal(tr.stats, ref.stats)
    finally:
        _lib.set_threads_per_chain(0)
