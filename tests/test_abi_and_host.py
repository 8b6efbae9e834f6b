"""CPU tests: the C-ABI library loads and exports every symbol the header declares;
host-side mirror of nutpie's settings/sampler surface behaves like the reference."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    txt = (ROOT / "include" / "nutpie_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nb200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from nutpie_b200 import _lib

    L = _lib.load_library()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"missing export {n}"
    assert L.nb200_abi_version() == 4


def test_settings_struct_layout_matches_header():
    from nutpie_b200 import _lib
    from oracle import pyoracle as O

    assert C.sizeof(_lib.Settings) == C.sizeof(O.Settings) == 240
    assert C.sizeof(_lib.ModelDesc) == 144
    s = _lib.Settings()
    _lib.load_library().nb200_settings_default(C.byref(s))
    d = O.default_settings()
    for name, _ in _lib.Settings._fields_:
        if name.startswith("_"):
            continue
        assert getattr(s, name) == getattr(d, name), name
    # SURVEY.md Appendix A.1 pins
    assert (s.num_tune, s.num_draws, s.maxdepth, s.target_accept) == (400, 1000, 10, 0.8)
    assert s.max_energy_error == 1000.0 and s.use_grad_based_estimate == 1


def test_settings_facade_follows_wrapper_rs():
    from nutpie_b200 import _lib

    s = _lib.PyNutsSettings.Diag(7)
    assert s.seed == 7
    s.update({"num_tune": 50, "num_draws": 20, "num_chains": 3, "maxdepth": 6, "target_accept": 0.9})
    assert (s.num_tune, s.num_draws, s.num_chains) == (50, 20, 3)
    s.use_grad_based_mass_matrix = False
    assert s._c.use_grad_based_estimate == 0
    s.step_size_adapt_method = "0.25"
    assert s._c.step_size_method == 2 and s._c.fixed_step_size == 0.25
    with pytest.raises(AttributeError):  # src/wrapper.rs:611-613
        s.no_such_option = 1
    with pytest.raises(ValueError):      # src/wrapper.rs:138-145
        s.mass_matrix_gamma = 1e-3
    with pytest.raises(ValueError):
        s.step_size_adapt_method = "bogus"
    d = s.as_dict()
    assert d["sampler"] == "nuts" and d["adaptation"] == "diag" and d["settings"]["maxdepth"] == 6
    assert _lib.PyNutsSettings.Diag(None).seed != _lib.PyNutsSettings.Diag(None).seed
    s.step_size_adapt_method = "adam"    # src/wrapper.rs:347-375
    s.step_size_adam_learning_rate = 0.1
    s.step_size_jitter = 0.2
    s.store_divergences = True
    assert (s._c.step_size_method, s._c.adam_learning_rate, s._c.step_size_jitter,
            s._c.store_divergences) == (1, 0.1, 0.2, 1)
    s.step_size_jitter = None
    assert s._c.step_size_jitter == 0.0
    lr = _lib.PyNutsSettings.LowRank(1)
    lr.mass_matrix_gamma = 1e-4
    assert lr.as_dict()["adaptation"] == "low_rank" and lr._c.mass_matrix_gamma == 1e-4
    with pytest.raises(ValueError):      # diag-only option on the low-rank settings
        lr.use_grad_based_mass_matrix = False


def test_sampler_fails_loudly_without_gpu():
    """No CPU fallback: on a box without a CUDA device creating a sampler raises."""
    import nutpie_b200
    from nutpie_b200 import _lib

    if _lib.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nutpie_b200.sample(nutpie_b200.normal_model(1), chains=2, draws=5, tune=5, seed=1)


def test_invalid_arguments_rejected():
    from nutpie_b200 import _lib

    L = _lib.load_library()
    s = _lib.Settings()
    L.nb200_settings_default(C.byref(s))
    d = _lib.ModelDesc()
    d.kind, d.dim = 99, 3
    assert not L.nb200_sampler_create(C.byref(s), C.byref(d), 1, 0, 0, None, None)
    assert b"unknown model" in L.nb200_last_error()
    d.kind, d.dim = 2, 1  # funnel needs dim >= 2
    assert not L.nb200_sampler_create(C.byref(s), C.byref(d), 1, 0, 0, None, None)
    d.kind, d.dim, d.sigma = 1, 1, 1.0
    s.maxdepth = 40
    assert not L.nb200_sampler_create(C.byref(s), C.byref(d), 1, 0, 0, None, None)
    assert b"maxdepth" in L.nb200_last_error()


def test_radon_layout_invariants(radon_data):
    """The host 'compile' step of the radon density: every observation lands in exactly
    one thread range, runs of a county are contiguous (models.cuh relies on it)."""
    d = radon_data
    order = np.argsort(d["county"], kind="stable")
    for T in (1, 32, 128):
        per = -(-len(order) // T)
        runs = []
        for t in range(T):
            seg = d["county"][order[t * per:(t + 1) * per]]
            if len(seg):
                runs += list(seg[np.r_[True, seg[1:] != seg[:-1]]])
        runs = np.array(runs)
        assert (np.diff(runs) >= 0).all()
        assert len(runs) <= d["n_county"] + T


def test_expand_matches_oracle_expand(radon_data):
    """CompiledDeviceModel._expand == oracle_expand_radon (src/pymc.rs:217-286)."""
    import nutpie_b200
    from oracle import pyoracle as O

    d = radon_data
    J = d["n_county"]; D = 2 * J + 5
    cm = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], J)
    om = O.Model("radon", D, y=d["y"], county=d["county"], floor=d["floor"], n_county=J)
    q = np.random.default_rng(0).normal(size=D) * 0.3
    out = np.zeros(4 * J + 5)
    rc = O.lib().oracle_expand_radon(C.c_size_t(D), C.c_size_t(4 * J + 5), q.ctypes.data_as(C.c_void_p),
                                     out.ctypes.data_as(C.c_void_p), om.ud_ptr)
    assert rc == 0
    ex = cm._expand(q)
    flat = np.concatenate([np.atleast_1d(ex[k]) for k in
                           ("intercept", "county_raw", "county_sd", "floor_effect", "county_floor_raw",
                            "county_floor_sd", "sigma", "county_effect", "county_floor_effect")])
    np.testing.assert_allclose(flat, out, rtol=1e-15)


def test_ess_estimator():
    from nutpie_b200.diagnostics import ess

    rng = np.random.default_rng(0)
    x = rng.normal(size=(8, 1000, 2))
    e = ess(x, None, device="cpu")
    assert (e > 6000).all() and (e < 10500).all()


def test_with_data_wrong_shape_raises_runtime_error(radon_data):
    """tests/test_pymc.py:418-420: wrong-shape with_data surfaces as RuntimeError."""
    import nutpie_b200

    d = radon_data
    cm = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], d["n_county"])
    ok = cm.with_data(y=d["y"] + 1.0)
    assert ok.params["y"][0] == d["y"][0] + 1.0 and cm.params["y"][0] == d["y"][0]
    with pytest.raises(RuntimeError):
        cm.with_data(y=d["y"][:-1])
    with pytest.raises(KeyError):
        cm.with_data(nope=np.zeros(3))


def test_unsupported_samplers_and_adaptations_are_explicit():
    """Out-of-scope variants fail loudly instead of silently running something else."""
    import nutpie_b200

    m = nutpie_b200.normal_model(2)
    for kw in (dict(adaptation="flow"), dict(sampler="mclmc")):
        with pytest.raises(NotImplementedError):
            nutpie_b200.sample(m, chains=1, draws=1, tune=1, **kw)
    with pytest.raises(ValueError):
        nutpie_b200.sample(m, chains=1, draws=1, tune=1, adaptation="bogus")
    with pytest.raises(NotImplementedError):
        nutpie_b200.compile_stan_model(code="parameters { real x; } model { x ~ normal(0, 1); }")


def test_expanded_layout_covers_every_variable(radon_data):
    import nutpie_b200

    d = radon_data
    J = d["n_county"]
    cm = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], J)
    lay = cm.expanded_layout()
    assert lay[0][1] == 0 and lay[-1][2] == 4 * J + 5
    for (_, a, b, shape), (_, a2, _, _) in zip(lay, lay[1:]):
        assert b == a2 and (b - a) == (int(np.prod(shape)) if shape else 1)
    assert {n for n, *_ in lay} == set(cm.dims)


# ------------------------------------------------- run-time compiled densities (no GPU needed)
def test_model_desc_layout():
    from nutpie_b200 import _lib

    assert C.sizeof(_lib.ModelDesc) == 144
    assert _lib.ModelDesc.cuda_source.offset == 64 and _lib.ModelDesc.n_user_scratch.offset == 88


@pytest.mark.parametrize("tpc,nit", [(32, 1), (32, 0), (256, 0)])
def test_custom_density_compiles_for_sm100a(tpc, nit):
    """NVRTC builds the sampler kernel around each test density for several geometries."""
    import nutpie_b200
    from nutpie_b200 import _lib
    from tests import custom_densities as CD

    for src, kw in ((CD.NORMAL, dict(data=[0.0, 1.0])), (CD.FUNNEL, {}),
                    (CD.LOGREG, dict(data=CD.logreg_data(20, 5), scratch=20)), (CD.WALL, {})):
        assert _lib.compile_custom(nutpie_b200.from_cuda_source(5, src, **kw), tpc, nit)


def test_custom_density_compile_errors_carry_the_nvrtc_log():
    import nutpie_b200
    from nutpie_b200 import _lib
    from tests import custom_densities as CD

    bad = nutpie_b200.from_cuda_source(5, CD.WALL.replace("acc += q[i] * q[i];", "acc += q[i] * nope;"))
    with pytest.raises(RuntimeError, match=r"nb200_user_density\.cu\(\d+\): error: identifier \"nope\""):
        _lib.compile_custom(bad, 32, 0)
    with pytest.raises(ValueError):   # no unrolled kernel with 5 dims per thread
        _lib.compile_custom(nutpie_b200.from_cuda_source(5, CD.WALL), 32, 5)
    with pytest.raises(ValueError):   # variable shapes must partition the coordinates
        nutpie_b200.from_cuda_source(5, CD.WALL, shapes={"a": (2,), "b": (2,)})
    with pytest.raises(ValueError):
        _lib.compile_custom(nutpie_b200.normal_model(3), 32, 0)


def test_custom_model_layout_and_with_data():
    import nutpie_b200
    from tests import custom_densities as CD

    m = nutpie_b200.from_cuda_source(12, CD.LOGREG, data=CD.logreg_data(20, 12), scratch=20,
                                     shapes={"intercept": (), "beta": (11,)})
    assert m.expanded_layout() == [("intercept", 0, 1, ()), ("beta", 1, 12, (11,))]
    e = np.arange(24.0).reshape(2, 12)
    out = m._split_expanded(e)
    assert out["intercept"].shape == (2,) and out["beta"].shape == (2, 11)
    m2 = m.with_data(data=CD.logreg_data(20, 12, seed=9))
    assert m2.kind == "custom" and not np.array_equal(m2.params["data"], m.params["data"])
    d, keep = m2._descriptor()
    assert d.kind == 4 and d.n_user_data == 2 + 20 * 12 + 20 and d.n_user_scratch == 20


def test_oracle_logreg_gradient_by_finite_differences():
    from oracle import pyoracle as O
    from tests import custom_densities as CD

    data = CD.logreg_data(50, 6)
    om = O.Model("logreg", 6, data=data)
    rng = np.random.default_rng(0)
    q = rng.normal(size=6)
    lp, g, rc = om.logp_grad(q)
    assert rc == 0
    for i in range(6):
        h = 1e-6
        e = np.zeros(6); e[i] = h
        fd = (om.logp_grad(q + e)[0] - om.logp_grad(q - e)[0]) / (2 * h)
        assert abs(fd - g[i]) < 1e-6 * max(1.0, abs(g[i]))


def test_arrow_batches_round_trip_through_the_reference_consumer():
    """`PyTrace.get_arrow_trace()` (src/wrapper.rs:1477-1494) hands one
    (posterior, sample_stats) RecordBatch pair per chain with `dims` / `shape` field metadata;
    `_arrow_to_groups` — `_arrow_to_arviz` / `_add_arrow_data` of python/nutpie/sample.py:62-214
    on the pyarrow API — must rebuild exactly what the dense path (`_trace_to_groups`) builds,
    including ragged chains (a run aborted mid-way) padded with NaN / 0."""
    pytest.importorskip("pyarrow")
    import nutpie_b200
    from nutpie_b200 import _lib
    import importlib

    S = importlib.import_module("nutpie_b200.sample")  # (the package re-exports the function `sample`)

    J, n_chains, tune, draws = 5, 3, 4, 6
    rng = np.random.default_rng(0)
    cm = nutpie_b200.radon_model(np.zeros(7), np.arange(7) % J, np.zeros(7), J)
    n_rows = tune + draws
    q = rng.normal(size=(n_chains, n_rows, cm.n_dim))
    stats = rng.normal(size=(n_chains, n_rows, _lib.NSTAT))
    stats[..., _lib.STAT_NAMES.index("tuning")] = (np.arange(n_rows) < tune)[None, :]
    for nm in ("depth", "n_steps", "draw", "chain"):
        stats[..., _lib.STAT_NAMES.index(nm)] = rng.integers(0, 9, size=(n_chains, n_rows))
    for nm in ("diverging", "maxdepth_reached"):
        stats[..., _lib.STAT_NAMES.index(nm)] = rng.integers(0, 2, size=(n_chains, n_rows))
    stats[..., _lib.STAT_NAMES.index("index_in_trajectory")] = rng.integers(-5, 5, size=(n_chains, n_rows))
    grads = rng.normal(size=(n_chains, n_rows, cm.n_dim))

    def trace(rows):
        return _lib.PyTrace(q, stats, np.asarray(rows, dtype=np.uint64), gradients=grads,
                            variables=cm._variable_dims(), expand=cm._expand)

    class _S:  # the settings fields _trace_to_groups reads
        num_tune = tune

        def as_dict(self):
            return {}

    # complete run: identical to the dense path, variable by variable
    dense = S._trace_to_groups(trace([n_rows] * n_chains), cm, _S(), True)
    post, st = trace([n_rows] * n_chains).get_arrow_trace()  # (draw batches, stat batches)
    assert len(post) == len(st) == n_chains
    assert post[0].schema.field("county_effect").metadata == {b"dims": b"county", b"shape": str(J).encode()}
    via = S._arrow_to_groups(post, st, skip_vars=["tuning", "draw", "chain"], coords=cm.coords)
    assert set(via.posterior) == set(dense.posterior)
    for name, a in dense.posterior.items():
        np.testing.assert_array_equal(via.posterior[name], a)
        np.testing.assert_array_equal(via.warmup_posterior[name], dense.warmup_posterior[name])
        assert via.dims[name] == list(cm.dims[name])
    for name, a in dense.sample_stats.items():
        np.testing.assert_array_equal(via.sample_stats[name], a)
        assert via.sample_stats[name].dtype == a.dtype, name
        np.testing.assert_array_equal(via.warmup_sample_stats[name], dense.warmup_sample_stats[name])
    assert via.dims["gradient"] == ["unconstrained_parameter"]
    assert "tuning" not in via.sample_stats
    # reparameterized value variables move out of the posterior (compile_pymc.py:810-814)
    uc = S._arrow_to_groups(post, st, reparameterized_names=["sigma"], keep_unconstrained_draw=True)
    assert "sigma" not in uc.posterior and uc.unconstrained_posterior["sigma"].shape == (n_chains, draws)
    # ragged chains: chain 1 stopped inside warm-up, chain 2 after two posterior draws
    rows = [n_rows, 3, tune + 2]
    rag = S._arrow_to_groups(*trace(rows).get_arrow_trace())
    sig = rag.posterior["sigma"]
    assert sig.shape == (n_chains, draws) and rag.warmup_posterior["sigma"].shape == (n_chains, tune)
    assert np.isnan(sig[1]).all() and np.isnan(sig[2, 2:]).all() and not np.isnan(sig[2, :2]).any()
    assert np.isnan(rag.warmup_posterior["sigma"][1, 3:]).all()
    np.testing.assert_array_equal(sig[0], np.exp(q[0, tune:, 2 * J + 4]))
    assert rag.sample_stats["n_steps"][1].sum() == 0  # integer columns pad with 0
