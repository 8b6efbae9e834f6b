"""adaptation="low_rank" (/root/reference/src/wrapper.rs:307-346, reference tests
tests/test_pymc.py:116-146): the low-rank modified mass matrix.

CPU part (-m "not gpu"): the oracle's restatement (oracle/lowrank.c) against numpy / a 60-digit
evaluation, and the CUDA engine's own code (csrc/lowrank.cuh, a different factorisation route)
run on the host by tests/emul against both.  GPU part (-m gpu): the same through the C-ABI.

Tolerances: the estimate is a matrix geometric mean of two covariances regularised with
gamma = 1e-5; with fewer draws than dimensions their condition number is ~1/gamma and the
oracle's route (eigen-decompositions of B and B^1/2 A B^1/2, like nuts-rs' spd_mean) loses about
half of the digits, the engine's Cholesky / one-sided-Jacobi route far fewer — both are checked
against the 60-digit value: engine <= 1e-8, oracle <= 1e-4 relative in the operator M^-1.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as O
from tests.emul import pyemul as E

STAT_N_STEPS, STAT_STEP, STAT_DIV = 9, 7, 6


def gaussian_window(dim, n, seed, noise=0.1):
    """Draws from a correlated Gaussian with its (noisy) gradients."""
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(dim, dim))
    cov = a @ a.T / dim + 0.05 * np.eye(dim)
    x = rng.multivariate_normal(np.zeros(dim), cov, size=n)
    g = -np.linalg.solve(cov, x.T).T + noise * rng.normal(size=(n, dim))
    return x, g


def operator(stds, vals, vecs):
    dim = len(stds)
    return np.array([O.lowrank_velocity(stds, vals, vecs, e) for e in np.eye(dim)])


def emul_component(x, g, gamma, cutoff, max_rank, p=None, z=None):
    L = E.lib()
    n, dim = x.shape
    x, g = np.ascontiguousarray(x), np.ascontiguousarray(g)
    p = np.eye(dim) if p is None else np.ascontiguousarray(p, dtype=np.float64).reshape(-1, dim)
    z = np.zeros_like(p) if z is None else np.ascontiguousarray(z, dtype=np.float64).reshape(-1, dim)
    v, pm = np.zeros_like(p), np.zeros_like(p)
    stds, vals, vecs = np.zeros(dim), np.zeros(max_rank), np.zeros((max_rank, dim))
    k = C.c_uint64(0)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    L.emul_lowrank_component.restype = C.c_int
    rc = L.emul_lowrank_component(C.c_uint64(dim), C.c_uint64(n), ptr(x), ptr(g), C.c_double(gamma),
                                  C.c_double(cutoff), C.c_uint64(max_rank), C.c_uint64(len(p)), ptr(p),
                                  ptr(v), ptr(z), ptr(pm), ptr(stds), ptr(vals), ptr(vecs), C.byref(k))
    assert rc == 0
    kk = int(k.value)
    return dict(stds=stds, vals=vals[:kk], vecs=vecs[:kk], velocity=v, momentum=pm)


def numpy_reference(x, g, gamma, cutoff):
    """The definition, with scipy's matrix square roots (fine when well conditioned)."""
    import scipy.linalg as sl

    n, dim = x.shape
    s = np.sqrt(x.std(0) / g.std(0))
    xs = (x - x.mean(0)) / s / np.sqrt(n)
    gs = (g - g.mean(0)) * s / np.sqrt(n)
    cx, cg = xs.T @ xs + gamma * np.eye(dim), gs.T @ gs + gamma * np.eye(dim)
    gh = sl.sqrtm(cg).real
    ghi = np.linalg.inv(gh)
    sig = ghi @ sl.sqrtm(gh @ cx @ gh).real @ ghi
    w, v = np.linalg.eigh(sig)
    keep = (w > cutoff) | (w < 1 / cutoff)
    minv = np.diag(s) @ (np.eye(dim) + v[:, keep] @ np.diag(w[keep] - 1) @ v[:, keep].T) @ np.diag(s)
    return s, np.sort(w[keep]), minv


def exact_reference(x, g, gamma, cutoff):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 60
    n, dim = x.shape
    s = np.sqrt(x.std(0) / g.std(0))
    xs = (x - x.mean(0)) / s / np.sqrt(n)
    gs = (g - g.mean(0)) * s / np.sqrt(n)
    cx = mp.matrix((xs.T @ xs + gamma * np.eye(dim)).tolist())
    cg = mp.matrix((gs.T @ gs + gamma * np.eye(dim)).tolist())

    def fun(a, f):
        w, v = mp.eigsy(a)
        return v * mp.diag([f(t) for t in w]) * v.T

    gh, ghi = fun(cg, mp.sqrt), fun(cg, lambda t: 1 / mp.sqrt(t))
    sig = ghi * fun(gh * cx * gh, mp.sqrt) * ghi
    w, v = mp.eigsy(sig)
    w = np.array([float(t) for t in w])
    v = np.array(v.tolist(), dtype=float)
    keep = (w > cutoff) | (w < 1 / cutoff)
    return np.diag(s) @ (np.eye(dim) + v[:, keep] @ np.diag(w[keep] - 1) @ v[:, keep].T) @ np.diag(s)


# ------------------------------------------------------------------ oracle components (CPU)
@pytest.mark.parametrize("cutoff", [2.0, 1.01])
def test_oracle_update_matches_the_definition(cutoff):
    x, g = gaussian_window(12, 40, seed=1, noise=0.0)
    stds, vals, vecs = O.lowrank_update(x, g, gamma=1e-5, cutoff=cutoff, max_rank=12)
    s, w, minv = numpy_reference(x, g, 1e-5, cutoff)
    np.testing.assert_allclose(stds, s, rtol=1e-14)
    np.testing.assert_allclose(np.sort(vals), w, rtol=1e-10)
    np.testing.assert_allclose(vecs @ vecs.T, np.eye(len(vals)), atol=1e-12)  # orthonormal
    np.testing.assert_allclose(operator(stds, vals, vecs), minv, rtol=0, atol=1e-10 * np.abs(minv).max())


def test_oracle_momentum_has_covariance_of_the_mass_matrix():
    """p = M^1/2 z: (M^1/2)(M^1/2)^T M^-1 = I, checked column by column."""
    x, g = gaussian_window(10, 30, seed=2)
    stds, vals, vecs = O.lowrank_update(x, g, cutoff=1.5, max_rank=10)
    assert len(vals) >= 3
    half = np.array([O.lowrank_momentum(stds, vals, vecs, e) for e in np.eye(10)]).T  # M^1/2
    minv = operator(stds, vals, vecs)
    np.testing.assert_allclose(half @ half.T @ minv, np.eye(10), atol=1e-10)


def test_oracle_max_rank_keeps_the_most_extreme_eigenvalues():
    x, g = gaussian_window(12, 40, seed=3)
    _, all_vals, _ = O.lowrank_update(x, g, cutoff=1.2, max_rank=12)
    _, vals, _ = O.lowrank_update(x, g, cutoff=1.2, max_rank=3)
    assert len(all_vals) > 3 and len(vals) == 3
    top = all_vals[np.argsort(-np.abs(np.log(all_vals)))[:3]]
    np.testing.assert_allclose(np.sort(vals), np.sort(top), rtol=1e-12)


def test_oracle_window_without_spread_keeps_the_previous_scale():
    x, g = gaussian_window(4, 10, seed=4)
    x[:, 2] = 0.5  # a coordinate that never moved (sums of 0.5 are exact: zero spread)
    stds, vals, vecs = O.lowrank_update(x, g, stds0=[1.0, 1.0, 3.5, 1.0], max_rank=4)
    assert stds[2] == 3.5 and np.all(np.isfinite(stds)) and np.all(np.isfinite(vals))


def test_jacobi_tournament_pairs_every_two_columns_exactly_once():
    """lr_round_robin_pair: m - 1 rounds of m / 2 disjoint pairs cover all m (m - 1) / 2 pairs."""
    L = E.lib()
    for m in (2, 4, 6, 12, 38, 176):
        seen = set()
        for t in range(m - 1):
            a, b = (C.c_int * (m // 2))(), (C.c_int * (m // 2))()
            L.emul_round_robin_pairs(C.c_int(m), C.c_int(t), a, b)
            cols = list(a) + list(b)
            assert sorted(cols) == list(range(m)), (m, t)  # a round touches every column once
            seen |= {frozenset(p) for p in zip(a, b)}
        assert len(seen) == m * (m - 1) // 2


# ------------------------------------------------- the engine's route on the host (tests/emul)
# (the last three take the engine's SUBSPACE route: fewer than 0.3 dim draws)
@pytest.mark.parametrize("dim,n,gamma", [(13, 11, 1e-5), (13, 40, 1e-5), (30, 20, 1e-5), (30, 20, 1e-3),
                                         (1, 12, 1e-5), (2, 3, 1e-5), (40, 8, 1e-5), (64, 12, 1e-5),
                                         (50, 15, 1e-3)])
def test_engine_and_oracle_routes_against_60_digits(dim, n, gamma):
    x, g = gaussian_window(dim, n, seed=5 + dim + n)
    exact = exact_reference(x, g, gamma, 2.0)
    scale = np.abs(exact).max()
    stds, vals, vecs = O.lowrank_update(x, g, gamma=gamma, cutoff=2.0, max_rank=dim)
    dev = emul_component(x, g, gamma, 2.0, dim)
    assert len(dev["vals"]) == len(vals)
    np.testing.assert_allclose(dev["stds"], stds, rtol=1e-13)
    assert np.abs(operator(stds, vals, vecs) - exact).max() <= 1e-4 * scale
    assert np.abs(dev["velocity"] - exact).max() <= 1e-8 * scale
    np.testing.assert_allclose(np.sort(dev["vals"]), np.sort(vals), rtol=1e-5)
    np.testing.assert_allclose(dev["vecs"] @ dev["vecs"].T, np.eye(len(vals)), atol=1e-10)


def test_engine_momentum_and_velocity_match_the_oracle_formulas():
    """Same metric in, same vectors out (to rounding): M^-1 p and M^1/2 z."""
    x, g = gaussian_window(9, 60, seed=8)
    rng = np.random.default_rng(0)
    p, z = rng.normal(size=(5, 9)), rng.normal(size=(5, 9))
    dev = emul_component(x, g, 1e-5, 2.0, 9, p=p, z=z)
    for i in range(5):
        np.testing.assert_allclose(dev["velocity"][i], O.lowrank_velocity(dev["stds"], dev["vals"], dev["vecs"], p[i]),
                                   rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(dev["momentum"][i], O.lowrank_momentum(dev["stds"], dev["vals"], dev["vecs"], z[i]),
                                   rtol=1e-12, atol=1e-14)


def lr_settings(**kw):
    base = dict(seed=11, num_tune=300, num_draws=100, adaptation=1, store_mass_matrix=1,
                mass_matrix_update_freq=10, mass_matrix_eigval_cutoff=2.0)
    base.update(kw)
    return O.default_settings(**base)


def small_radon(J=4, N=60, seed=0):
    rng = np.random.default_rng(seed)
    return dict(y=rng.normal(1.0, 0.8, N), county=rng.integers(0, J, N).astype(np.int32),
                floor=rng.integers(0, 2, N).astype(np.uint8), n_county=J)


@pytest.mark.parametrize("dim", [5, 40])
def test_emulated_engine_equals_oracle_on_an_isotropic_density(dim):
    """iid normal: whichever eigenpairs a noisy window throws up, both implementations must take
    the same decisions — trees, window switches, refresh schedule, step-size search.  dim = 40: the
    early windows (10-20 draws) go through the engine's subspace route."""
    s = lr_settings(num_tune=150 if dim == 40 else 300, num_draws=50 if dim == 40 else 100)
    ro = O.sample(O.Model("normal", dim, mu=1.0, sigma=2.0), s, 4)
    re = E.sample("normal", dim, s, 4, mu=1.0, sigma=2.0)
    same = ro["stats"][..., STAT_N_STEPS] == re["stats"][..., STAT_N_STEPS]
    assert same[:, :12].all() and same.mean() > (0.9 if dim == 40 else 0.999)
    np.testing.assert_allclose(re["draws"][:, :11], ro["draws"][:, :11], rtol=0, atol=1e-9)
    np.testing.assert_allclose(re["mass_matrix_inv"][:, :12], ro["mass_matrix_inv"][:, :12], rtol=1e-9)


@pytest.mark.parametrize("kind", ["funnel", "radon"])
def test_emulated_engine_follows_oracle_until_the_first_refresh_then_statistically(kind):
    kw = small_radon() if kind == "radon" else {}
    dim = 13 if kind == "radon" else 9
    s = lr_settings(num_tune=400, num_draws=300)
    ro = O.sample(O.Model(kind, dim, **kw), s, 6)
    re = E.sample(kind, dim, s, 6, **kw)
    # identity metric until the refresh after draw 10: same trees, positions to rounding
    assert np.array_equal(ro["stats"][:, :11, STAT_N_STEPS], re["stats"][:, :11, STAT_N_STEPS])
    np.testing.assert_allclose(re["draws"][:, :11], ro["draws"][:, :11], rtol=0, atol=1e-8)
    # the refreshed scales agree to the accuracy of the two routes
    np.testing.assert_allclose(re["mass_matrix_inv"][:, 11], ro["mass_matrix_inv"][:, 11], rtol=1e-6)
    po, pe = ro["stats"][:, 400:], re["stats"][:, 400:]
    assert abs(po[..., STAT_STEP].mean() / pe[..., STAT_STEP].mean() - 1) < 0.25
    assert abs(po[..., STAT_N_STEPS].mean() / pe[..., STAT_N_STEPS].mean() - 1) < 0.35
    if kind == "radon":
        do, de = ro["draws"][:, 400:], re["draws"][:, 400:]
        sd = do.std((0, 1))
        assert np.all(np.abs(do.mean((0, 1)) - de.mean((0, 1))) < 0.35 * sd)


def test_emulated_engine_resumes_bit_identically_in_low_rank_mode():
    """Chunked relaunches: the window deque, rank and metric persist between launches."""
    s = lr_settings(num_tune=120, num_draws=30)
    kw = small_radon()
    a = E.sample("radon", 13, s, 2, **kw)
    b = E.sample("radon", 13, s, 2, max_per_launch=7, **kw)
    assert np.array_equal(a["draws"], b["draws"]) and np.array_equal(a["stats"], b["stats"])


@pytest.mark.parametrize("kind,dim,tune", [("radon", 13, 100), ("normal", 40, 50)])
def test_lane_emulation_of_the_low_rank_engine_is_schedule_independent(kind, dim, tune):
    """The low-rank engine run by 32 cooperative lanes (tests/emul GroupLanes: the lanes run one
    after another between two barriers — in ascending, descending or freshly shuffled order, all
    of them schedules independent thread scheduling allows): identical traces under every order,
    i.e. no barrier is missing in lowrank.cuh / leapfrog_lr; and the same trees as the one-thread
    emulation over the first draws.  normal-40: the early windows take the subspace route."""
    kw = small_radon() if kind == "radon" else dict(mu=1.0, sigma=2.0)
    s = lr_settings(num_tune=tune, num_draws=10)
    runs = [E.sample_lanes(kind, dim, s, 1, threads_per_chain=32, lane_order=o, **kw) for o in (0, 1, 5)]
    for other in runs[1:]:
        assert np.array_equal(other["draws"], runs[0]["draws"]) and np.array_equal(other["stats"], runs[0]["stats"])
    serial = E.sample(kind, dim, s, 1, **kw)
    assert np.array_equal(serial["stats"][:, :11, STAT_N_STEPS], runs[0]["stats"][:, :11, STAT_N_STEPS])
    np.testing.assert_allclose(runs[0]["draws"][:, :11], serial["draws"][:, :11], rtol=0, atol=1e-8)


# ------------------------------------------------------------------- oracle sampler (CPU)
def test_oracle_low_rank_shortens_trajectories_on_a_correlated_posterior():
    rng = np.random.default_rng(3)
    n, d = 300, 8
    zz = rng.normal(size=(n, d))
    zz[:, 1] = zz[:, 0] * 0.95 + 0.1 * zz[:, 1]
    zz[:, 3] = zz[:, 2] * 0.9 + 0.2 * zz[:, 3]
    beta = rng.normal(size=d)
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-zz @ beta))).astype(float)
    m = O.Model("logreg", d, data=np.concatenate([[n, d], zz.ravel(), y]))
    out = {}
    for adapt in (0, 1):
        s = O.default_settings(seed=5, num_tune=400, num_draws=400, adaptation=adapt, store_mass_matrix=1,
                               mass_matrix_update_freq=10 if adapt else 1)
        out[adapt] = O.sample(m, s, 8)
    diag, low = out[0], out[1]
    assert low["stats"][:, 400:, STAT_N_STEPS].mean() < 0.5 * diag["stats"][:, 400:, STAT_N_STEPS].mean()
    assert low["stats"][:, 400:, STAT_DIV].sum() == 0
    dm, lm = diag["draws"][:, 400:], low["draws"][:, 400:]
    assert np.all(np.abs(dm.mean((0, 1)) - lm.mean((0, 1))) < 0.15 * dm.std((0, 1)))
    assert np.all(np.abs(dm.std((0, 1)) / lm.std((0, 1)) - 1) < 0.1)
    eig = low["mass_matrix_eigvals"][:, -1]
    assert np.isfinite(eig).sum(axis=1).min() >= 2  # the two correlated pairs are found
    assert diag["mass_matrix_eigvals"] is None


# ----------------------------------------------------------------------------- GPU (-m gpu)
@pytest.mark.gpu
# (300: more than 256 rows — the two-pass Jacobi loop instead of the register-resident one)
@pytest.mark.parametrize("dim,n", [(13, 11), (30, 20), (64, 200), (175, 90), (64, 12), (175, 20), (175, 50), (300, 100)])
def test_gpu_refresh_matches_its_host_emulation_and_the_oracle(dim, n):
    from nutpie_b200 import _lib

    x, g = gaussian_window(dim, n, seed=dim)
    rng = np.random.default_rng(1)
    p, z = rng.normal(size=(4, dim)), rng.normal(size=(4, dim))
    dev = _lib.lowrank_component(x, g, gamma=1e-5, cutoff=2.0, max_rank=32, p=p, z=z)
    emu = emul_component(x, g, 1e-5, 2.0, min(32, dim), p=p, z=z)
    assert len(dev["vals"]) == len(emu["vals"])
    np.testing.assert_allclose(dev["stds"], emu["stds"], rtol=1e-12)
    # (device and host run the same algorithm with different roundings — FMA contraction, rsqrt —
    # and the problem amplifies them: both are within 1e-8 of the 60-digit value on the sizes where
    # that is computable, test_engine_and_oracle_routes_against_60_digits)
    np.testing.assert_allclose(np.sort(dev["vals"]), np.sort(emu["vals"]), rtol=2e-6)
    scale = np.abs(emu["velocity"]).max()
    np.testing.assert_allclose(dev["velocity"], emu["velocity"], rtol=0, atol=1e-6 * scale)
    np.testing.assert_allclose(dev["momentum"], emu["momentum"], rtol=0, atol=1e-6 * np.abs(emu["momentum"]).max())
    if dim <= 64:  # the oracle's dense Jacobi route (accuracy ~1e-6 when rank deficient)
        stds, vals, vecs = O.lowrank_update(x, g, max_rank=min(32, dim))
        ref = np.array([O.lowrank_velocity(stds, vals, vecs, v) for v in p])
        np.testing.assert_allclose(dev["velocity"], ref, rtol=0, atol=2e-5 * scale)


def _gpu_pair(seed=11, **kw):
    from nutpie_b200 import _lib

    s = _lib.PyNutsSettings.LowRank(seed)
    so = lr_settings(seed=seed)
    base = dict(num_tune=300, num_draws=100, store_mass_matrix=1, mass_matrix_update_freq=10,
                mass_matrix_eigval_cutoff=2.0)
    base.update(kw)
    for k, v in base.items():
        setattr(s._c, k, v)
        setattr(so, k, v)
    return s, so


def _run_gpu(s, model, n_chains, **kw):
    from nutpie_b200 import _lib

    smp = _lib.PySampler(s, model, n_chains=n_chains, **kw)
    try:
        smp.wait()
        return smp.take_results(), smp.geometry()
    finally:
        smp.close()


@pytest.mark.gpu
def test_gpu_low_rank_equals_oracle_on_an_isotropic_density():
    import nutpie_b200

    s, so = _gpu_pair()
    tr, geom = _run_gpu(s, nutpie_b200.normal_model(5, 1.0, 2.0), 6)
    assert geom["threads_per_chain"] == 32
    ref = O.sample(O.Model("normal", 5, mu=1.0, sigma=2.0), so, 6)
    same = tr.stats[..., STAT_N_STEPS] == ref["stats"][..., STAT_N_STEPS]
    assert same[:, :40].all() and same.mean() > 0.97
    np.testing.assert_allclose(tr.draws[:, :12], ref["draws"][:, :12], rtol=0, atol=1e-9)
    assert tr.low_rank and tr.mass_matrix_eigvals.shape == (6, 400, 5)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["funnel", "radon_small", "radon"])
def test_gpu_low_rank_follows_oracle_then_agrees_statistically(kind, radon_data):
    import nutpie_b200

    if kind == "funnel":
        gm, om, chains = nutpie_b200.funnel_model(9), O.Model("funnel", 9), 64
    elif kind == "radon_small":
        kw = small_radon()
        gm = nutpie_b200.radon_model(kw["y"], kw["county"], kw["floor"], kw["n_county"])
        om, chains = O.Model("radon", 13, **kw), 64
    else:
        d = radon_data
        J = d["n_county"]
        gm = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], J)
        om = O.Model("radon", 2 * J + 5, y=d["y"], county=d["county"], floor=d["floor"], n_county=J)
        chains = 8
    tune = 300 if kind == "radon" else 400
    s, so = _gpu_pair(num_tune=tune, num_draws=200, mass_matrix_eigval_cutoff=4.0 if kind == "radon" else 2.0)
    tr, _ = _run_gpu(s, gm, chains)
    ref = O.sample(om, so, chains)
    # (rounding differences grow along the 255-leapfrog trajectories of the first, badly scaled draws)
    assert np.array_equal(tr.stats[:, :11, STAT_N_STEPS], ref["stats"][:, :11, STAT_N_STEPS])
    np.testing.assert_allclose(tr.draws[:, :4], ref["draws"][:, :4], rtol=0, atol=1e-8)
    np.testing.assert_allclose(tr.draws[:, :11], ref["draws"][:, :11], rtol=0, atol=2e-2)
    np.testing.assert_allclose(tr.mass_matrix_inv[:, 11], ref["mass_matrix_inv"][:, 11], rtol=2e-2)
    pg, po = tr.stats[:, tune:], ref["stats"][:, tune:]
    assert abs(pg[..., STAT_STEP].mean() / po[..., STAT_STEP].mean() - 1) < 0.2
    assert abs(pg[..., STAT_N_STEPS].mean() / po[..., STAT_N_STEPS].mean() - 1) < 0.3
    dg, do = tr.draws[:, tune:], ref["draws"][:, tune:]
    sd = do.std((0, 1))
    tol = 0.5 if kind == "radon" else 0.3
    if kind != "funnel":
        assert np.all(np.abs(dg.mean((0, 1)) - do.mean((0, 1))) < tol * sd)


@pytest.mark.gpu
def test_gpu_low_rank_pause_resume_and_chunked_launches_are_bit_identical():
    import nutpie_b200

    kw = small_radon()
    gm = nutpie_b200.radon_model(kw["y"], kw["county"], kw["floor"], kw["n_county"])
    mk = lambda: _gpu_pair(num_tune=150, num_draws=50)[0]
    a, _ = _run_gpu(mk(), gm, 12)
    b, _ = _run_gpu(mk(), gm, 12, draws_per_launch=13)
    assert np.array_equal(a.draws, b.draws) and np.array_equal(a.stats, b.stats)
    assert np.array_equal(np.isnan(a.mass_matrix_eigvals), np.isnan(b.mass_matrix_eigvals))


@pytest.mark.gpu
def test_gpu_low_rank_through_sample_like_the_reference_tests():
    """/root/reference/tests/test_pymc.py:116-146: one parameter; 'mass_matrix_eigvals' is a stat
    only with store_mass_matrix; a 45-dimensional model samples."""
    import nutpie_b200

    m = nutpie_b200.normal_model(1)
    tr = nutpie_b200.sample(m, chains=1, adaptation="low_rank", progress_bar=False, seed=1)
    assert "mass_matrix_eigvals" not in tr.sample_stats
    assert tr.posterior["x"].shape[:2] == (1, 1000)
    tr = nutpie_b200.sample(m, chains=1, adaptation="low_rank", store_mass_matrix=True, progress_bar=False, seed=1)
    assert "mass_matrix_eigvals" in tr.sample_stats and "mass_matrix_stds" in tr.sample_stats
    assert "mass_matrix_inv" not in tr.sample_stats
    x = tr.posterior["x"]
    assert abs(x.mean()) < 0.15 and abs(x.std() - 1) < 0.1
    tr = nutpie_b200.sample(nutpie_b200.normal_model(45, 0.5, 3.0), chains=2, adaptation="low_rank",
                            mass_matrix_eigval_cutoff=3, mass_matrix_gamma=1e-5, progress_bar=False, seed=2)
    x = tr.posterior["x"]
    assert abs(x.mean() - 0.5) < 0.2 and abs(x.std() - 3.0) < 0.2
    with pytest.raises(ValueError):  # low-rank options are rejected by the diagonal adaptation
        nutpie_b200.sample(m, chains=1, mass_matrix_gamma=1e-3, progress_bar=False)


@pytest.mark.gpu
def test_gpu_low_rank_shortens_trajectories_on_a_correlated_run_time_compiled_density():
    """A correlated logistic regression as user CUDA source (NB200_MODEL_CUSTOM: the low-rank engine
    is compiled by NVRTC around the user's density): the metric finds the correlated directions,
    needs far fewer leapfrogs than the diagonal one, and both agree with the oracle's posterior."""
    import nutpie_b200
    from nutpie_b200 import _lib
    from tests import custom_densities as CD

    rng = np.random.default_rng(3)
    n, d = 300, 8
    zz = rng.normal(size=(n, d))
    zz[:, 1] = zz[:, 0] * 0.95 + 0.1 * zz[:, 1]
    zz[:, 3] = zz[:, 2] * 0.9 + 0.2 * zz[:, 3]
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-zz @ rng.normal(size=d)))).astype(float)
    flat = np.concatenate([[n, d], zz.ravel(), y])
    gm = nutpie_b200.from_cuda_source(d, CD.LOGREG, data=flat, scratch=n)
    res = {}
    for name, mk in (("diag", lambda: _lib.PyNutsSettings.Diag(5)), ("low_rank", lambda: _lib.PyNutsSettings.LowRank(5))):
        s = mk()
        s._c.num_tune, s._c.num_draws = 400, 300
        res[name], _ = _run_gpu(s, gm, 32)
    nd = res["diag"].stats[:, 400:, STAT_N_STEPS].mean()
    nl = res["low_rank"].stats[:, 400:, STAT_N_STEPS].mean()
    assert nl < 0.6 * nd
    dm, lm = res["diag"].draws[:, 400:], res["low_rank"].draws[:, 400:]
    assert np.all(np.abs(dm.mean((0, 1)) - lm.mean((0, 1))) < 0.15 * dm.std((0, 1)))
    so = O.default_settings(seed=5, num_tune=400, num_draws=300, adaptation=1, mass_matrix_update_freq=10)
    ref = O.sample(O.Model("logreg", d, data=flat), so, 32)
    om = ref["draws"][:, 400:]
    assert np.all(np.abs(om.mean((0, 1)) - lm.mean((0, 1))) < 0.15 * om.std((0, 1)))
    assert abs(ref["stats"][:, 400:, STAT_N_STEPS].mean() / nl - 1) < 0.25


@pytest.mark.gpu
def test_gpu_low_rank_through_the_reference_plugin_abi(radon_data):
    """adaptation="low_rank" with the density behind the reference's HOST plug-in ABI
    (src/pymc.rs:23-29): same posterior as the hand-written device density."""
    import nutpie_b200
    from tests.test_host_plugin import _radon_pointer_model

    d = radon_data
    cm_host, _ = _radon_pointer_model(d)
    cm_dev = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], d["n_county"])
    kw = dict(chains=8, tune=200, draws=100, seed=3, adaptation="low_rank", mass_matrix_eigval_cutoff=4.0,
              progress_bar=False)
    a, b = nutpie_b200.sample(cm_host, **kw), nutpie_b200.sample(cm_dev, **kw)
    for name in ("intercept", "sigma"):
        xa, xb = a.posterior[name], b.posterior[name]
        assert abs(xa.mean() - xb.mean()) < 0.5 * xb.std(), name


@pytest.mark.gpu
def test_gpu_deprecated_keywords_and_extra_stats_like_the_reference_tests(radon_data):
    """/root/reference/tests/test_pymc.py:150-173 (deprecated `low_rank_modified_mass_matrix` /
    `use_grad_based_mass_matrix` still work and warn) and :305-327 (all four store_* flags at once)."""
    import nutpie_b200

    m = nutpie_b200.normal_model(1)
    with pytest.warns(FutureWarning, match="low_rank_modified_mass_matrix"):
        tr = nutpie_b200.sample(m, chains=1, low_rank_modified_mass_matrix=True, progress_bar=False, seed=3)
    assert tr.posterior["x"].shape[:2] == (1, 1000)
    with pytest.warns(FutureWarning, match="use_grad_based_mass_matrix"):
        tr = nutpie_b200.sample(m, chains=1, use_grad_based_mass_matrix=False, progress_bar=False, seed=3)
    assert tr.posterior["x"].shape[:2] == (1, 1000)
    d = radon_data
    rm = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], d["n_county"])
    for adaptation in ("diag", "low_rank"):
        tr = nutpie_b200.sample(rm, chains=2, tune=60, draws=40, seed=4, adaptation=adaptation, progress_bar=False,
                                store_mass_matrix=True, store_divergences=True, store_unconstrained=True,
                                store_gradient=True)
        ss = tr.sample_stats
        for name in ("gradient", "unconstrained_draw", "divergence_start", "divergence_momentum"):
            assert ss[name].shape[:2] == (2, 40) and ss[name].shape[-1] == rm.n_dim, name
        mm = "mass_matrix_stds" if adaptation == "low_rank" else "mass_matrix_inv"
        assert ss[mm].shape == (2, 40, rm.n_dim) and np.isfinite(ss[mm]).all()
        assert ("mass_matrix_eigvals" in ss) == (adaptation == "low_rank")


def test_low_rank_settings_are_validated_by_the_c_abi_before_any_device_work():
    """nb200_sampler_create refuses what the low-rank engine cannot run (no GPU needed: the checks
    come before the first CUDA call), with the option names of src/wrapper.rs:307-346."""
    import nutpie_b200
    from nutpie_b200 import _lib

    def make(dim=3, **kw):
        s = _lib.PyNutsSettings.LowRank(1)
        for k, v in kw.items():
            setattr(s._c, k, v)
        return _lib.PySampler(s, nutpie_b200.normal_model(dim), n_chains=2, autostart=False)

    with pytest.raises(ValueError, match="1024 dimensions"):
        make(dim=2000)
    with pytest.raises(ValueError, match="mass_matrix_eigval_cutoff"):
        make(mass_matrix_eigval_cutoff=1.0)
    with pytest.raises(ValueError, match="mass_matrix_gamma"):
        make(mass_matrix_gamma=0.0)
    with pytest.raises(ValueError, match="mass_matrix_max_rank"):
        make(mass_matrix_max_rank=0)
    s = _lib.PyNutsSettings.LowRank(1)
    assert (s._c.adaptation, s._c.mass_matrix_update_freq, s.num_tune) == (1, 10, 800)
    assert s.as_dict()["adaptation"] == "low_rank"
    with pytest.raises(ValueError):  # wrapper.rs:138-145: a diag-only option on the low-rank settings
        s.use_grad_based_mass_matrix = False
