"""GPU parity tests (-m gpu): the CUDA engine, called through the C-ABI, against the CPU
oracle on identical seeds/inputs.

Tolerances (fp64, BASELINE.json north_star "stated fp64 tolerance"):
  * single density / leapfrog evaluations: 1e-12 relative (reduction order + FMA only);
  * whole runs on order-independent densities (normal): identical tree shapes
    (depth, n_steps, index_in_trajectory, diverging) and positions within 1e-9 over the first draws and 1e-3 after 500 draws (rounding
    differences — FMA contraction, reduction order, libm — are amplified by the early
    mass-matrix estimates, which divide small sums of squares, and grow along a run);
  * radon / funnel (chaotic amplification of rounding differences): identical for the
    first draws, then statistical parity — means within 4 MCSE, sds within 5 %,
    step size within 10 % (SURVEY.md §8d "Parity report").
"""
import json

import numpy as np
import pytest

import nutpie_b200
from nutpie_b200 import _lib
from oracle import pyoracle as O

from tests import custom_densities as CD

pytestmark = pytest.mark.gpu
LOGREG_DATA = CD.logreg_data(200, 12)

STAT = {n: i for i, n in enumerate(_lib.STAT_NAMES)}


DEFAULT_STAGE_MODE = 3  # bulk-copy staging + alternating sweep direction (nb200_set_stage_loads)


def settings_pair(seed=1, **kw):
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
        return smp.take_results()
    finally:
        smp.close()


@pytest.fixture(autouse=True)
def _reset_geometry():
    yield
    _lib.set_threads_per_chain(0)
    _lib.set_chains_per_block(0)
    _lib.set_smem_slots(-1)
    _lib.set_stage_loads(DEFAULT_STAGE_MODE)


def models(radon_data):
    d = radon_data
    J = d["n_county"]
    return {
        "normal1": (nutpie_b200.normal_model(1), O.Model("normal", 1)),
        "normal37": (nutpie_b200.normal_model(37, 2.0, 0.5), O.Model("normal", 37, mu=2.0, sigma=0.5)),
        "funnel": (nutpie_b200.funnel_model(9), O.Model("funnel", 9)),
        "radon": (nutpie_b200.radon_model(d["y"], d["county"], d["floor"], J),
                  O.Model("radon", 2 * J + 5, y=d["y"], county=d["county"], floor=d["floor"], n_county=J)),
        # run-time compiled CUDA densities (NB200_MODEL_CUSTOM) against their host twins
        "c_normal37": (nutpie_b200.from_cuda_source(37, CD.NORMAL, data=[2.0, 4.0]),
                       O.Model("normal", 37, mu=2.0, sigma=0.5)),
        "c_funnel": (nutpie_b200.from_cuda_source(9, CD.FUNNEL), O.Model("funnel", 9)),
        "c_logreg": (nutpie_b200.from_cuda_source(12, CD.LOGREG, data=LOGREG_DATA, scratch=200),
                     O.Model("logreg", 12, data=LOGREG_DATA)),
    }


@pytest.mark.parametrize("tpc", [32, 128, 1024])
@pytest.mark.parametrize("name", ["normal1", "normal37", "funnel", "radon", "c_normal37", "c_funnel",
                                  "c_logreg"])
def test_logp_grad_matches_oracle(radon_data, name, tpc):
    gm, om = models(radon_data)[name]
    _lib.set_threads_per_chain(tpc)
    rng = np.random.default_rng(0)
    q = rng.normal(size=(24, gm.n_dim)) * 0.6
    lp, g, rc = _lib.logp_grad(gm, q)
    lpo, go, rco = om.logp_grad(q)
    np.testing.assert_allclose(lp, lpo, rtol=1e-12)
    np.testing.assert_allclose(g, go, rtol=1e-11, atol=1e-11 * np.abs(go).max())
    assert (rc == 0).all() and (rco == 0).all()


def test_logp_nonfinite_codes():
    gm = nutpie_b200.funnel_model(3)
    lp, g, rc = _lib.logp_grad(gm, np.array([[-800.0, 1.0, 1.0], [0.1, 1.0, 1.0]]))
    assert rc[0] in (3, 4) and rc[1] == 0


@pytest.mark.parametrize("tpc", [32, 256])
@pytest.mark.parametrize("name", ["normal37", "funnel", "radon", "c_funnel", "c_logreg"])
def test_leapfrog_matches_oracle(radon_data, name, tpc):
    import ctypes as C

    gm, om = models(radon_data)[name]
    _lib.set_threads_per_chain(tpc)
    D = gm.n_dim
    rng = np.random.default_rng(1)
    n = 12
    q = rng.normal(size=(n, D)) * 0.4
    p = rng.normal(size=(n, D))
    var = np.exp(rng.normal(size=(n, D)) * 0.5)
    psum = rng.normal(size=(n, D))
    _, g, _ = om.logp_grad(q)
    eps = rng.uniform(0.01, 0.1, size=n)
    direction = np.where(rng.random(n) < 0.5, 1, -1).astype(np.int32)
    idx = np.where(direction > 0, rng.integers(0, 5, n), -rng.integers(0, 5, n)).astype(np.int64)
    out = _lib.leapfrog(gm, q, p, g, var, psum, eps, direction, idx)
    L = O.lib()
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    for i in range(n):
        o = [np.zeros(D) for _ in range(4)]
        lp, kin = C.c_double(), C.c_double()
        L.oracle_leapfrog(om.fn_ptr, om.ud_ptr, C.c_size_t(D), ptr(q[i]), ptr(p[i]), ptr(g[i]), ptr(var[i]),
                          ptr(psum[i]), C.c_double(eps[i]), C.c_int(int(direction[i])), C.c_int64(int(idx[i])),
                          ptr(o[0]), ptr(o[1]), ptr(o[2]), ptr(o[3]), C.byref(lp), C.byref(kin))
        for key, ref in zip(("q", "p", "g", "p_sum"), o):
            np.testing.assert_allclose(out[key][i], ref, rtol=1e-11, atol=1e-11 * max(1.0, np.abs(ref).max()), err_msg=key)
        np.testing.assert_allclose(out["logp"][i], lp.value, rtol=1e-12)
        np.testing.assert_allclose(out["kinetic"][i], kin.value, rtol=1e-12)


@pytest.mark.parametrize("tpc,slots", [(32, -1), (32, 0), (64, -1), (256, 2)])
@pytest.mark.parametrize("name", ["normal1", "normal37"])
def test_sampler_matches_oracle_draw_for_draw(radon_data, name, tpc, slots):
    """Identical seeds -> identical tree shapes and positions to rounding, including
    warm-up adaptation (step size, mass matrix)."""
    gm, om = models(radon_data)[name]
    _lib.set_threads_per_chain(tpc)
    _lib.set_smem_slots(slots)
    s, so = settings_pair(seed=5, num_tune=300, num_draws=200, store_mass_matrix=1, store_gradient=1)
    tr = run_gpu(s, gm, 16)
    ref = O.sample(om, so, 16)
    for k in ("depth", "n_steps", "index_in_trajectory", "diverging", "maxdepth_reached", "tuning"):
        assert np.array_equal(tr.stats[..., STAT[k]], ref["stats"][..., STAT[k]]), k
    # rounding differences (FMA contraction, reduction order, libm) grow along a run:
    # tight on the first draws, loose at the end of 500 draws
    np.testing.assert_allclose(tr.draws[:, :3], ref["draws"][:, :3], rtol=0, atol=1e-9)
    np.testing.assert_allclose(tr.draws, ref["draws"], rtol=0, atol=1e-3)
    np.testing.assert_allclose(tr.stats[..., STAT["step_size"]], ref["stats"][..., STAT["step_size"]], rtol=1e-2)
    np.testing.assert_allclose(tr.stats[..., STAT["step_size_bar"]], ref["stats"][..., STAT["step_size_bar"]], rtol=1e-2)
    np.testing.assert_allclose(tr.mass_matrix_inv, ref["mass_matrix_inv"], rtol=1e-2)
    np.testing.assert_allclose(tr.stats[..., STAT["energy"]], ref["stats"][..., STAT["energy"]], rtol=1e-3, atol=1e-2)


def test_sampler_tape_driven_fixed_step(radon_data):
    """same z tape, fixed step size, no adaptation."""
    gm, om = models(radon_data)["normal37"]
    rng = np.random.default_rng(3)
    z = rng.normal(size=(4, 120, 37))
    q0 = rng.normal(size=(4, 37)) * 0.5 + 2.0
    s, so = settings_pair(seed=2, num_tune=0, num_draws=120, step_size_method=2, fixed_step_size=0.2)
    tr = run_gpu(s, gm, 4, q0=q0, z_tape=z)
    ref = O.sample(om, so, 4, q0=q0, z_tape=z)
    for k in ("depth", "n_steps", "index_in_trajectory"):
        assert np.array_equal(tr.stats[..., STAT[k]], ref["stats"][..., STAT[k]]), k
    np.testing.assert_allclose(tr.draws, ref["draws"], atol=1e-10)
    np.testing.assert_allclose(tr.stats[..., STAT["logp"]], ref["stats"][..., STAT["logp"]], rtol=1e-10)


def _mcse_compare(a, b):
    """a, b: [chain, draw, dim]; compare means in units of combined MCSE (chains as
    independent replicates) and sds relatively."""
    ma, mb = a.mean(1), b.mean(1)  # per-chain means
    se = np.sqrt(ma.var(0, ddof=1) / len(ma) + mb.var(0, ddof=1) / len(mb))
    zscore = np.abs(ma.mean(0) - mb.mean(0)) / se
    sd_ratio = a.std((0, 1)) / b.std((0, 1))
    return zscore, sd_ratio


def test_radon_parity(radon_data):
    gm, om = models(radon_data)["radon"]
    s, so = settings_pair(seed=7, num_tune=400, num_draws=400, init_radius=1.0)
    n = 96
    tr = run_gpu(s, gm, n)
    ref = O.sample(om, so, n)
    # identical start: first draws agree to rounding before chaos separates the runs
    dd = np.abs(tr.draws - ref["draws"]).max(axis=(0, 2))
    assert dd[0] < 1e-9 and dd[:4].max() < 1e-6
    z, r = _mcse_compare(tr.draws[:, 400:], ref["draws"][:, 400:])
    assert z.max() < 4.5, z.max()
    assert np.abs(r - 1).max() < 0.05, r
    g_step = tr.stats[:, -1, STAT["step_size"]].mean()
    o_step = ref["stats"][:, -1, STAT["step_size"]].mean()
    assert abs(g_step / o_step - 1) < 0.10
    g_n = tr.stats[:, 400:, STAT["n_steps"]].mean()
    o_n = ref["stats"][:, 400:, STAT["n_steps"]].mean()
    assert abs(g_n / o_n - 1) < 0.10
    assert abs(tr.stats[:, 400:, STAT["mean_tree_accept"]].mean() - 0.8) < 0.06
    assert tr.stats[:, 400:, STAT["diverging"]].mean() < 0.01


def test_funnel_divergence_and_depth_parity(radon_data):
    """BASELINE config 5 (reduced chains): divergence rate and depth histogram vs oracle."""
    gm, om = models(radon_data)["funnel"]
    s, so = settings_pair(seed=11, num_tune=300, num_draws=300, maxdepth=12)
    n = 256
    tr = run_gpu(s, gm, n)
    ref = O.sample(om, so, n)
    gd = tr.stats[:, 300:, STAT["diverging"]].mean()
    od = ref["stats"][:, 300:, STAT["diverging"]].mean()
    assert abs(gd - od) < 0.01 + 0.25 * od, (gd, od)
    hg = np.bincount(tr.stats[:, 300:, STAT["depth"]].astype(int).ravel(), minlength=13) / (n * 300)
    ho = np.bincount(ref["stats"][:, 300:, STAT["depth"]].astype(int).ravel(), minlength=13) / (n * 300)
    assert np.abs(hg - ho).max() < 0.03, (hg, ho)
    # log_sigma marginal ~ N(0,1) up to funnel bias, same in both
    assert abs(tr.draws[:, 300:, 0].mean() - ref["draws"][:, 300:, 0].mean()) < 0.1


def test_determinism_and_seed_contract():
    """tests/test_stan.py:67-101, 282-302: same seed -> bit-identical (also on the GPU,
    no atomics in the data path); different seed / different chain -> different."""
    m = nutpie_b200.funnel_model(9)
    a = run_gpu(settings_pair(seed=42, num_tune=100, num_draws=50)[0], m, 32)
    b = run_gpu(settings_pair(seed=42, num_tune=100, num_draws=50)[0], m, 32)
    c = run_gpu(settings_pair(seed=43, num_tune=100, num_draws=50)[0], m, 32)
    assert np.array_equal(a.draws, b.draws) and np.array_equal(a.stats, b.stats)
    assert not np.array_equal(a.draws, c.draws)
    for i in range(1, 32):
        assert not np.allclose(a.draws[0], a.draws[i])


def test_radon_bitwise_reproducible(radon_data):
    gm, _ = models(radon_data)["radon"]
    for tpc in (32, 128):
        _lib.set_threads_per_chain(tpc)
        a = run_gpu(settings_pair(seed=1, num_tune=60, num_draws=40, init_radius=1.0)[0], gm, 24)
        b = run_gpu(settings_pair(seed=1, num_tune=60, num_draws=40, init_radius=1.0)[0], gm, 24)
        assert np.array_equal(a.draws, b.draws)


def test_chain_sharding_reproduces_single_run(radon_data):
    """RNG keyed by global chain id -> a run split over several samplers
    (GPUs) equals the unsplit run chain for chain."""
    gm, _ = models(radon_data)["radon"]
    s = lambda: settings_pair(seed=9, num_tune=80, num_draws=40, init_radius=1.0)[0]
    full = run_gpu(s(), gm, 16)
    lo = run_gpu(s(), gm, 8)
    hi = run_gpu(s(), gm, 8, chain_id_offset=8)
    assert np.array_equal(full.draws[:8], lo.draws) and np.array_equal(full.draws[8:], hi.draws)
    assert np.array_equal(full.stats[8:, :, STAT["chain"]], hi.stats[:, :, STAT["chain"]])


def test_pause_resume_and_chunked_launch_are_bit_identical(radon_data):
    gm, _ = models(radon_data)["radon"]
    mk = lambda: settings_pair(seed=4, num_tune=150, num_draws=100, init_radius=1.0)[0]
    ref = run_gpu(mk(), gm, 16)
    chunked = run_gpu(mk(), gm, 16, draws_per_launch=37)
    assert np.array_equal(ref.draws, chunked.draws) and np.array_equal(ref.stats, chunked.stats)
    smp = _lib.PySampler(mk(), gm, n_chains=16)
    try:
        smp.pause()
        part = smp.inspect()
        assert part.rows_filled.max() <= 250
        smp.resume()
        smp.wait()
        tr = smp.take_results()
    finally:
        smp.close()
    assert np.array_equal(ref.draws, tr.draws)


def test_config4_leapfrog_properties():
    """BASELINE config 4 at full dimension (D = 10 000, fewer chains): size-independent
    properties — posterior sd ~ 1 per coordinate, energy error small, acceptance near
    the target, and the first draws equal to the oracle's."""
    D = 10000
    gm, om = nutpie_b200.normal_model(D), O.Model("normal", D)
    s, so = settings_pair(seed=3, num_tune=120, num_draws=60, store_dims=16)
    tr = run_gpu(s, gm, 12)
    ref = O.sample(om, so, 12)
    for k in ("depth", "n_steps", "index_in_trajectory"):
        assert np.array_equal(tr.stats[..., STAT[k]], ref["stats"][..., STAT[k]]), k
    np.testing.assert_allclose(tr.draws[:, :3], ref["draws"][:, :3], rtol=0, atol=1e-8)  # sums over 10^4 terms
    np.testing.assert_allclose(tr.draws, ref["draws"], rtol=0, atol=1e-3)
    post = tr.draws[:, 120:]
    assert abs(post.std() - 1.0) < 0.08
    assert np.abs(tr.stats[:, 120:, STAT["energy_error"]]).max() < 5.0
    assert abs(tr.stats[:, 120:, STAT["mean_tree_accept"]].mean() - 0.8) < 0.1
    # logp of a D-dim standard normal draw: -chi2_D/2 ~ -D/2 +- sqrt(D/2)
    assert abs(tr.stats[:, 120:, STAT["logp"]].mean() + D / 2) < 5 * np.sqrt(D / 2)


@pytest.mark.parametrize("tpc", [128, 256])
def test_streaming_leapfrog_staging_modes(tpc):
    """The streaming leapfrog of config 4 (bulk-copy staging through shared memory, 128 or 256
    threads per chain) on an odd, ragged dimension: every mode reproduces the oracle's tree
    shapes; staging on/off is bit-identical for the same sweep order, and L2 hints never change
    a bit."""
    D = 4099  # odd (scalar tail) and not a multiple of the chunk size
    gm, om = nutpie_b200.normal_model(D, mu=0.3, sigma=1.7), O.Model("normal", D, mu=0.3, sigma=1.7)
    mk = lambda: settings_pair(seed=5, num_tune=70, num_draws=30, store_dims=24)
    ref = O.sample(om, mk()[1], 6)
    _lib.set_threads_per_chain(tpc)
    out = {}
    for mode in (0, 1, 2, 3, 7):
        _lib.set_stage_loads(mode)
        tr = run_gpu(mk()[0], gm, 6)
        out[mode] = tr
        for k in ("depth", "n_steps", "index_in_trajectory", "diverging"):
            assert np.array_equal(tr.stats[..., STAT[k]], ref["stats"][..., STAT[k]]), (mode, k)
        np.testing.assert_allclose(tr.draws[:, :3], ref["draws"][:, :3], rtol=0, atol=1e-9, err_msg=str(mode))
        np.testing.assert_allclose(tr.draws, ref["draws"], rtol=0, atol=1e-4, err_msg=str(mode))  # rounding growth
        np.testing.assert_allclose(tr.stats[..., STAT["energy"]], ref["stats"][..., STAT["energy"]],
                                   rtol=0, atol=1e-3)
    for a, b in ((0, 1), (2, 3), (3, 7)):
        assert np.array_equal(out[a].draws, out[b].draws), (a, b)
        assert np.array_equal(out[a].stats, out[b].stats), (a, b)


def test_config1_through_public_api():
    """BASELINE config 1 through nutpie_b200.sample (the reference's public call)."""
    tr = nutpie_b200.sample(nutpie_b200.normal_model(1), chains=4, draws=1000, tune=400, seed=0,
                            progress_bar=False)
    assert set(tr.groups()) >= {"posterior", "sample_stats", "warmup_posterior", "warmup_sample_stats"}
    x = tr.posterior["x"]
    assert x.shape == (4, 1000) and tr.warmup_posterior["x"].shape == (4, 400)
    assert abs(x.mean()) < 0.1 and abs(x.std() - 1) < 0.06
    assert tr.sample_stats["diverging"].dtype == np.bool_
    assert tr.sample_stats["depth"].shape == (4, 1000)


def test_public_api_radon_shapes_and_save_warmup(radon_data):
    d = radon_data
    J = d["n_county"]
    cm = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], J)
    tr = nutpie_b200.sample(cm, chains=8, draws=50, tune=100, seed=1, save_warmup=False,
                            init_radius=1.0, store_gradient=True)
    assert tr.posterior["county_effect"].shape == (8, 50, J)
    assert tr.posterior["sigma"].shape == (8, 50) and (tr.posterior["sigma"] > 0).all()
    assert not tr.warmup_posterior
    assert tr.sample_stats["gradient"].shape == (8, 50, 2 * J + 5)
    np.testing.assert_allclose(tr.posterior["county_effect"],
                               tr.posterior["county_raw"] * tr.posterior["county_sd"][..., None])
    raw = nutpie_b200.sample(cm, chains=2, draws=10, tune=10, seed=1, return_raw_trace=True,
                             init_radius=1.0)
    draw_batches, stat_batches = raw.get_arrow_trace()  # src/wrapper.rs:1477-1494
    assert len(draw_batches) == len(stat_batches) == 2 and draw_batches[0].num_rows == 20
    assert stat_batches[0].column("tuning").to_numpy(zero_copy_only=False).sum() == 10
    with pytest.raises(ValueError):
        raw.get_arrow_trace()  # single-take, src/wrapper.rs:1477-1494


def test_nonblocking_control_surface(radon_data):
    """tests/test_pymc.py:224-286: non-blocking sampling, wait(timeout) raising
    TimeoutError, pause/resume, abort returning a partial trace."""
    big = nutpie_b200.normal_model(100_000)  # the reference uses a 100 000-dim model too
    bg = nutpie_b200.sample(big, chains=64, draws=2000, tune=2000, seed=1, blocking=False,
                            store_dims=4, progress_bar=False)
    with pytest.raises(TimeoutError):
        bg.wait(timeout=0.1)
    assert not bg.is_finished
    bg.pause()
    bg.resume()
    part = bg.abort()
    n = part.posterior["unconstrained_draw"].shape[1] + part.warmup_posterior["unconstrained_draw"].shape[1]
    assert 0 <= n < 4000
    bg.close()


def test_progress_callback_contents():
    """tests/test_pymc.py:37-66"""
    seen = []
    tr = nutpie_b200.sample(nutpie_b200.normal_model(50_000), chains=32, draws=300, tune=300, seed=2,
                            progress_callback=lambda p: seen.append(p), progress_rate=20,
                            store_dims=2)
    assert seen, "callback never fired"
    last = seen[-1]
    assert len(last) == 32
    p = last[0]
    assert p.total_draws == 600 and 0 <= p.finished_draws <= 600
    assert p.step_size > 0 and p.total_num_steps >= p.latest_num_steps
    assert isinstance(p.divergent_draws, list) and p.runtime_ms >= 0


def test_init_failure_is_reported():
    """A chain that finds no finite initial point surfaces as RuntimeError from wait
    (src/wrapper.rs:1131-1136), not as silent garbage."""
    m = nutpie_b200.funnel_model(3)
    s = _lib.PyNutsSettings.Diag(1)
    s.update({"num_tune": 10, "num_draws": 10})
    q0 = np.array([[-800.0, 1.0, 1.0]])
    smp = _lib.PySampler(s, m, n_chains=1, q0=q0)
    with pytest.raises(RuntimeError, match="initial point"):
        smp.wait()
    smp.close()


def test_streamed_trace_equals_final_copy(radon_data):
    """Rows streamed to (pinned) host buffers while the kernel runs == one copy at the end."""
    gm, _ = models(radon_data)["radon"]
    mk = lambda: settings_pair(seed=6, num_tune=300, num_draws=300, init_radius=1.0)[0]
    ref = run_gpu(mk(), gm, 64)
    n_rows, D = 600, gm.n_dim
    # the engine's buffers are row-major: [row][chain][width]
    pd_, ps_ = _lib.PinnedArray((n_rows, 64, D)), _lib.PinnedArray((n_rows, 64, _lib.NSTAT))
    pd_.array[:] = -1.0
    bufs = {"draws": pd_.array, "stats": ps_.array}
    smp = _lib.PySampler(mk(), gm, n_chains=64, trace_buffers=bufs)
    try:
        smp.wait()
        tr = smp.take_results()
    finally:
        smp.close()
    assert np.shares_memory(tr.draws, bufs["draws"]) and tr.draws.shape == (64, n_rows, D)
    assert np.array_equal(tr.draws, ref.draws) and np.array_equal(tr.stats, ref.stats)


def test_device_expand_matches_host_expand(radon_data):
    """expand_vector on the device (src/pymc.rs:217-286): the expanded trace equals the host
    expansion of the unconstrained trace of the same run, and the oracle's expand twin."""
    import ctypes as C

    gm, om = models(radon_data)["radon"]
    J = radon_data["n_county"]
    mk = lambda e: settings_pair(seed=8, num_tune=60, num_draws=40, init_radius=1.0, expand_draws=e)[0]
    raw = run_gpu(mk(0), gm, 8)
    exp_ = run_gpu(mk(1), gm, 8)
    assert exp_.expanded and exp_.draws.shape == (8, 100, 4 * J + 5)
    host = gm._expand(raw.draws)
    dev = gm._split_expanded(exp_.draws)
    for k in host:
        np.testing.assert_allclose(dev[k], host[k], rtol=1e-15, atol=0, err_msg=k)
    assert np.array_equal(exp_.stats, raw.stats)
    # oracle twin on one draw
    out = np.zeros(4 * J + 5)
    q = np.ascontiguousarray(raw.draws[3, 57])
    assert O.lib().oracle_expand_radon(C.c_size_t(2 * J + 5), C.c_size_t(4 * J + 5), q.ctypes.data_as(C.c_void_p),
                                       out.ctypes.data_as(C.c_void_p), om.ud_ptr) == 0
    np.testing.assert_allclose(exp_.draws[3, 57], out, rtol=1e-15)


def test_pymc_model_shared_pin_on_gpu():
    """tests/test_pymc.py:397-416 through the public API: posterior mean of N(-0.1, 1)^3
    within 0.05 and of N(10, 3)^3 within 0.5."""
    tr = nutpie_b200.sample(nutpie_b200.normal_model(3, -0.1, 1.0), chains=4, draws=1000, tune=400, seed=1)
    np.testing.assert_allclose(tr.posterior["x"].mean(), -0.1, atol=0.05)
    tr = nutpie_b200.sample(nutpie_b200.normal_model(3, 10.0, 3.0), chains=4, draws=1000, tune=400, seed=1)
    np.testing.assert_allclose(tr.posterior["x"].mean(), 10.0, atol=0.5)


def test_draw_diag_adaptation_and_var_names(radon_data):
    """adaptation="draw_diag" (python/nutpie/sample.py:1015-1023) and var_names filtering
    (tests/test_pymc.py:423-468)."""
    d = radon_data
    J = d["n_county"]
    cm = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], J)
    tr = nutpie_b200.sample(cm, chains=16, draws=200, tune=300, seed=3, adaptation="draw_diag",
                            init_radius=1.0, var_names=["sigma", "county_effect"])
    assert set(tr.posterior) == {"sigma", "county_effect"}
    assert tr.posterior["county_effect"].shape == (16, 200, J)
    assert abs(tr.posterior["sigma"].mean() - d["truth"]["sigma"]) < 0.1
    assert tr.dims["county_effect"] == ["county"] and len(tr.coords["county"]) == J
    assert json.loads(tr.attrs["_settings"])["settings"]["use_grad_based_estimate"] == 0


# ------------------------------------------------- run-time compiled densities (NVRTC)
@pytest.mark.parametrize("tpc", [0, 64])
def test_custom_normal_sampler_matches_oracle_draw_for_draw(radon_data, tpc):
    """The NVRTC-compiled sampler kernel around a user density reproduces the oracle run on the
    twin host density: identical tree shapes, positions to rounding."""
    gm, om = models(radon_data)["c_normal37"]
    _lib.set_threads_per_chain(tpc)
    s, so = settings_pair(seed=5, num_tune=300, num_draws=200)
    tr = run_gpu(s, gm, 16)
    ref = O.sample(om, so, 16)
    for k in ("depth", "n_steps", "index_in_trajectory", "diverging", "maxdepth_reached", "tuning"):
        assert np.array_equal(tr.stats[..., STAT[k]], ref["stats"][..., STAT[k]]), k
    np.testing.assert_allclose(tr.draws[:, :3], ref["draws"][:, :3], rtol=0, atol=1e-9)
    np.testing.assert_allclose(tr.draws, ref["draws"], rtol=0, atol=1e-3)


def test_custom_equals_builtin_density_bitwise_shapes(radon_data):
    """Same density, built-in elementwise kernel vs run-time compiled gather kernel."""
    gm, _ = models(radon_data)["c_normal37"]
    bm, _ = models(radon_data)["normal37"]
    s, _ = settings_pair(seed=11, num_tune=200, num_draws=100)
    a = run_gpu(s, gm, 32)
    b = run_gpu(s, bm, 32)
    for k in ("depth", "n_steps", "index_in_trajectory", "diverging"):
        assert np.array_equal(a.stats[..., STAT[k]], b.stats[..., STAT[k]]), k
    np.testing.assert_allclose(a.draws, b.draws, rtol=0, atol=1e-3)


def test_custom_logreg_parity(radon_data):
    """A user model with no built-in counterpart: logistic regression from CUDA source vs the
    oracle on the host twin — identical start, then statistical parity."""
    gm, om = models(radon_data)["c_logreg"]
    s, so = settings_pair(seed=2, num_tune=300, num_draws=300)
    C = 64
    tr = run_gpu(s, gm, C)
    ref = O.sample(om, so, C)
    np.testing.assert_allclose(tr.draws[:, 0], ref["draws"][:, 0], rtol=0, atol=1e-9)
    assert np.array_equal(tr.stats[:, :20, STAT["n_steps"]], ref["stats"][:, :20, STAT["n_steps"]])
    a, b = tr.draws[:, 300:], ref["draws"][:, 300:]
    z, ratio = _mcse_compare(a, b)
    assert z.max() < 4.5, z
    assert np.all(np.abs(ratio - 1) < 0.08), ratio
    sa = tr.stats[:, -1, STAT["step_size_bar"]].mean()
    sb = ref["stats"][:, -1, STAT["step_size_bar"]].mean()
    assert abs(sa / sb - 1) < 0.1
    dg, do = tr.stats[..., STAT["diverging"]], ref["stats"][..., STAT["diverging"]]
    assert dg[:, 300:].sum() == do[:, 300:].sum() == 0          # none after warm-up
    assert np.array_equal(dg[:, :10], do[:, :10])               # same early-warm-up divergences
    assert abs(dg.sum() - do.sum()) <= 0.2 * max(do.sum(), 10)


def test_custom_recoverable_error_is_a_divergence():
    """rc > 0 from the density is the reference's recoverable error (src/pymc.rs:178): the
    trajectory ends there as a divergence and no draw lands in the forbidden region."""
    gm = nutpie_b200.from_cuda_source(4, CD.WALL)
    s, _ = settings_pair(seed=3, num_tune=100, num_draws=200)
    tr = run_gpu(s, gm, 32, q0=np.full((32, 4), 0.2))
    assert tr.draws[..., 0].max() <= 1.0
    assert tr.stats[..., STAT["diverging"]].sum() > 0
    lp, g, rc = _lib.logp_grad(gm, np.array([[2.0, 0, 0, 0], [0.5, 0, 0, 0]]))
    assert rc[0] != 0 and rc[1] == 0


def test_custom_compile_error_surfaces_at_create():
    gm = nutpie_b200.from_cuda_source(4, CD.WALL.replace("acc += q[i] * q[i];", "acc += q[i] * nope;"))
    s, _ = settings_pair(seed=3, num_tune=10, num_draws=10)
    with pytest.raises(RuntimeError, match="nope"):
        run_gpu(s, gm, 4)


def test_custom_public_api_variables():
    tr = nutpie_b200.sample(
        nutpie_b200.from_cuda_source(12, CD.LOGREG, data=LOGREG_DATA, scratch=200,
                                     shapes={"intercept": (), "beta": (11,)}),
        chains=8, draws=50, tune=100, seed=4, progress_bar=False)
    post = tr.posterior
    assert np.asarray(post["intercept"]).shape == (8, 50)
    assert np.asarray(post["beta"]).shape == (8, 50, 11)


def test_single_process_multi_device_equals_one_device(radon_data):
    """ONE `sample(chains=...)` call over several devices of the process
    (one persistent kernel + one host thread per device, the analogue of `cores`,
    src/wrapper.rs:977-1085) equals the one-device run chain for chain.  With a single GPU on
    the box the two shards run as two concurrent samplers of device 0."""
    d = radon_data
    J = d["n_county"]
    cm = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], J)
    kw = dict(chains=20, draws=60, tune=90, seed=5, init_radius=1.0, progress_bar=False)
    one = nutpie_b200.sample(cm, **kw)
    devs = [0, 0, 0] if _lib.device_count() < 3 else [0, 1, 2]
    seen = []
    multi = nutpie_b200.sample(cm, devices=devs, progress_callback=lambda ps: seen.append(len(ps)),
                               progress_rate=10, **kw)
    for name in ("intercept", "county_effect", "sigma"):
        assert np.array_equal(one.posterior[name], multi.posterior[name]), name
        assert np.array_equal(one.warmup_posterior[name], multi.warmup_posterior[name]), name
    for name in ("n_steps", "step_size", "diverging"):
        assert np.array_equal(one.sample_stats[name], multi.sample_stats[name]), name
    assert not seen or set(seen) == {20}
    raw = nutpie_b200.sample(cm, devices=devs, return_raw_trace=True, **kw)
    assert isinstance(raw, _lib.MultiTrace) and [p.draws.shape[0] for p in raw.parts] == [7, 7, 6]
    assert raw.draws.shape[0] == 20 and list(raw.stats[:, 0, STAT["chain"]]) == list(range(20))


@pytest.mark.parametrize("dim,tpc", [(1, 4), (1, 8), (3, 4), (9, 4), (9, 8), (12, 16), (30, 16)])
def test_sub_warp_geometry_matches_oracle_draw_for_draw(dim, tpc):
    """Sub-warp groups (4 / 8 / 16 lanes per chain, 32 / lanes chains per warp: the geometry of
    BASELINE configs 1 and 5): chains that share a warp part ways wherever their trees differ,
    yet every chain reproduces the oracle's trees exactly on an order-independent density —
    and 70 chains (not a multiple of the chains per warp or per CTA) all finish."""
    # (sigma = 1/2: inv_var is a power of two, so FMA contraction on the device rounds like the oracle)
    gm, om = nutpie_b200.normal_model(dim, 2.0, 0.5), O.Model("normal", dim, mu=2.0, sigma=0.5)
    _lib.set_threads_per_chain(tpc)
    s, so = settings_pair(seed=8, num_tune=200, num_draws=100, store_mass_matrix=1)
    smp = _lib.PySampler(s, gm, n_chains=70)
    try:
        smp.wait()
        tr, geom = smp.take_results(), smp.geometry()
    finally:
        smp.close()
    assert geom["threads_per_chain"] == tpc
    ref = O.sample(om, so, 70)
    # the trees are the oracle's; a last-bit difference in a sum (lane-strided partial sums on
    # the device, a sequential loop in the oracle) flips a decision once in ~10^4 draws, after
    # which that chain follows its own path
    for k in ("depth", "n_steps", "index_in_trajectory", "diverging", "maxdepth_reached"):
        same = tr.stats[..., STAT[k]] == ref["stats"][..., STAT[k]]
        assert same[:, :30].all() and same.mean() > 0.98, (k, same.mean())
    np.testing.assert_allclose(tr.draws[:, :3], ref["draws"][:, :3], rtol=0, atol=1e-9)
    intact = (tr.stats[..., STAT["n_steps"]] == ref["stats"][..., STAT["n_steps"]]).all(axis=1)
    assert intact.mean() > 0.9
    np.testing.assert_allclose(tr.draws[intact], ref["draws"][intact], rtol=0, atol=1e-3)
    np.testing.assert_allclose(tr.mass_matrix_inv[intact], ref["mass_matrix_inv"][intact], rtol=1e-2)


def test_sub_warp_funnel_equals_warp_per_chain_at_first_then_statistically():
    """The gathering density path (shared-memory front, group reductions) on 8-lane groups: the
    first draws equal the warp-per-chain run to rounding; pause / resume and chunked launches
    stay bit-identical; auto geometry picks the sub-warp kernel for D = 9."""
    m = nutpie_b200.funnel_model(9)
    mk = lambda: settings_pair(seed=21, num_tune=150, num_draws=100, maxdepth=12)[0]
    _lib.set_threads_per_chain(32)
    warp = run_gpu(mk(), m, 40)
    _lib.set_threads_per_chain(0)
    smp = _lib.PySampler(mk(), m, n_chains=40)
    try:
        smp.wait()
        sub, geom = smp.take_results(), smp.geometry()
    finally:
        smp.close()
    assert geom["threads_per_chain"] == 8
    np.testing.assert_allclose(sub.draws[:, :3], warp.draws[:, :3], rtol=1e-9, atol=1e-11)
    assert np.mean(sub.stats[:, :30, STAT["n_steps"]] == warp.stats[:, :30, STAT["n_steps"]]) > 0.9
    chunked = run_gpu(mk(), m, 40, draws_per_launch=23)
    assert np.array_equal(chunked.draws, sub.draws) and np.array_equal(chunked.stats, sub.stats)


def test_two_warp_pipeline_is_bit_identical(radon_data):
    """The opt-in producer / consumer kernel (an integrator warp and a tree warp per chain,
    profiles/r2_pipeline_notes.txt) computes the same arithmetic in the same order as the
    one-warp kernel: identical traces, also across pause / chunked relaunches."""
    gm, _ = models(radon_data)["radon"]
    mk = lambda: settings_pair(seed=31, num_tune=120, num_draws=80, init_radius=1.0, store_divergences=1)[0]
    ref = run_gpu(mk(), gm, 20)
    _lib.set_pipeline(True)
    try:
        smp = _lib.PySampler(mk(), gm, n_chains=20)
        try:
            smp.wait()
            tr, geom = smp.take_results(), smp.geometry()
        finally:
            smp.close()
        assert geom["pipelined"]
        chunked = run_gpu(mk(), gm, 20, draws_per_launch=17)
    finally:
        _lib.set_pipeline(False)
    for other in (tr, chunked):
        assert np.array_equal(other.draws, ref.draws) and np.array_equal(other.stats, ref.stats)
    assert np.array_equal(np.isnan(tr.divergences), np.isnan(ref.divergences))


@pytest.mark.parametrize("tpc", [0, 64])
def test_adam_step_size_and_jitter_match_oracle(radon_data, tpc):
    """step_size_adapt_method="adam" + step_size_jitter (src/wrapper.rs:347-407): same trees as
    the oracle on an order-independent density, same step-size trajectory."""
    gm, om = models(radon_data)["normal37"]
    _lib.set_threads_per_chain(tpc)
    s, so = settings_pair(seed=13, num_tune=250, num_draws=100, step_size_method=1,
                          adam_learning_rate=0.1, step_size_jitter=0.3)
    tr = run_gpu(s, gm, 12)
    ref = O.sample(om, so, 12)
    for k in ("depth", "n_steps", "index_in_trajectory", "diverging"):
        assert np.array_equal(tr.stats[..., STAT[k]], ref["stats"][..., STAT[k]]), k
    np.testing.assert_allclose(tr.stats[..., STAT["step_size"]], ref["stats"][..., STAT["step_size"]], rtol=1e-6)
    ss = tr.stats[:, 250:, STAT["step_size"]]
    assert ss.std(axis=1).min() > 0.05 * ss.mean()          # jittered after tuning as well
    assert abs(tr.stats[:, 250:, STAT["mean_tree_accept"]].mean() - 0.8) < 0.1


def test_store_divergences_rows_match_oracle(radon_data):
    """store_divergences (src/wrapper.rs:438-442; python/nutpie/sample.py:641-646): start / end
    location, start momentum and start gradient of the diverging leapfrog, NaN rows otherwise.
    A tight max_energy_error makes an order-independent density 'diverge' often."""
    gm, om = models(radon_data)["normal37"]
    s, so = settings_pair(seed=17, num_tune=60, num_draws=60, max_energy_error=0.4, store_divergences=1)
    tr = run_gpu(s, gm, 8)
    ref = O.sample(om, so, 8)
    assert np.array_equal(tr.stats[..., STAT["diverging"]], ref["stats"][..., STAT["diverging"]])
    div = tr.stats[..., STAT["diverging"]] > 0
    assert 5 < div.sum() < div.size
    assert tr.divergences.shape == (8, 120, 4, 37)
    assert np.isnan(tr.divergences[~div]).all() and np.isfinite(tr.divergences[div]).all()
    np.testing.assert_allclose(tr.divergences[div], ref["divergences"][div], rtol=1e-6, atol=1e-8)
    res = nutpie_b200.sample(gm, chains=4, tune=40, draws=40, seed=17, max_energy_error=0.4,
                             store_divergences=True, progress_bar=False)
    for name in _lib.DIVERGENCE_COLUMNS:
        assert res.sample_stats[name].shape == (4, 40, 37)


def test_trace_lands_identically_through_every_host_path(radon_data):
    """The trace of one job through the three device -> host paths of nb200_api.cu: rows streamed
    into pageable arrays (the default: pinned staging ring + parallel host copies, several chunks
    per block), rows streamed into caller-provided pinned buffers (direct DMA), and one copy
    when the trace is taken (trace_buffers=False; an 80 MB block = two staging chunks)."""
    gm, _ = models(radon_data)["radon"]
    mk = lambda: settings_pair(seed=41, num_tune=150, num_draws=150, init_radius=1.0)[0]
    default = run_gpu(mk(), gm, 192)
    single = run_gpu(mk(), gm, 192, trace_buffers=False)
    pd_, ps_ = _lib.PinnedArray((300, 192, 175)), _lib.PinnedArray((300, 192, _lib.NSTAT))
    pinned = run_gpu(mk(), gm, 192, trace_buffers={"draws": pd_.array, "stats": ps_.array})
    assert default.draws.shape == (192, 300, 175) and np.isfinite(default.draws).all()
    for other in (single, pinned):
        assert np.array_equal(other.draws, default.draws) and np.array_equal(other.stats, default.stats)


def test_shards_of_one_job_stream_into_one_result_array(radon_data):
    """sample(devices=...) / PyMultiSampler: every shard writes its chains into its own block of
    columns of the job's single [row][chain][width] array (row-strided trace targets,
    nb200_sampler_set_trace_target_strided) — pageable (staging ring, row-wise host copies) and
    pinned (one 2-D DMA per block).  Two shards on ONE device here; chain for chain the
    single-sampler run, and the result is a view of one array, not a concatenation."""
    gm, _ = models(radon_data)["radon"]
    mk = lambda: settings_pair(seed=43, num_tune=100, num_draws=60, init_radius=1.0)[0]
    single = run_gpu(mk(), gm, 40)
    ms = _lib.PyMultiSampler(mk(), gm, n_chains=40, devices=[0, 0])
    try:
        ms.wait()
        tr = ms.take_results()
        big = ms._big["draws"]
    finally:
        ms.close()
    assert np.shares_memory(tr.draws, big) and tr.draws.shape == (40, 160, 175)
    assert np.array_equal(tr.draws, single.draws) and np.array_equal(tr.stats, single.stats)
    pd_, ps_ = _lib.PinnedArray((160, 40, 175)), _lib.PinnedArray((160, 40, _lib.NSTAT))
    bufs = [{"draws": pd_.array[:, :20], "stats": ps_.array[:, :20]},
            {"draws": pd_.array[:, 20:], "stats": ps_.array[:, 20:]}]
    ms = _lib.PyMultiSampler(mk(), gm, n_chains=40, devices=[0, 0], trace_buffers=bufs)
    try:
        ms.wait()
        ms.take_results()
    finally:
        ms.close()
    assert np.array_equal(pd_.array.transpose(1, 0, 2), single.draws)
    assert np.array_equal(ps_.array.transpose(1, 0, 2), single.stats)
