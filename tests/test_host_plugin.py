"""The reference's own density plug-in ABI honoured by the engine.

`int logp(size_t dim, const double* x, double* grad, double* logp, const void* user_data)`
(src/pymc.rs:23-29; producer python/nutpie/compile_pymc.py:970-1006) is sampled through
NB200_MODEL_HOST: the persistent CUDA kernel posts positions to a mailbox in mapped pinned
memory and host threads call the pointer.  The tests build the same objects the reference's
Python layer builds — LogpFunc / ExpandFunc / PyVariable.new_variables / PyMcModel
(compile_pymc.py:189-233), PyModel (compiled_pyfunc.py:72-105) — with a PyMC-free fixture: a
numba @cfunc, the oracle's `oracle_logp_radon` pointer, and Python callables."""
import ctypes as C

import numpy as np
import pytest

import nutpie_b200
from nutpie_b200 import _lib, compile as NC
from oracle import pyoracle as O


# ----------------------------------------------------------------------------- CPU: b1 classes
def test_new_variables_follows_common_rs():
    """src/common.rs:302-465: consecutive slices, anonymous dims generated and written back
    into the caller's dicts, known dims give the shape, inconsistent sizes are refused."""
    dim_sizes, dims = {"county": 85}, {"a": ["county"]}
    vs = _lib.PyVariable.new_variables(["mu", "a", "b"], ["float64", "float64", "int64"],
                                       [[], None, [2, 3]], dim_sizes, dims)
    assert [(v.name, v.start_idx, v.end_idx, v.num_elements) for v in vs] == \
        [("mu", 0, 1, 1), ("a", 1, 86, 85), ("b", 86, 92, 6)]
    assert vs[1].shape == [85] and vs[1].dims == ["county"]
    assert vs[2].dims == ["b_dim_0", "b_dim_1"]
    assert dim_sizes == {"county": 85, "b_dim_0": 2, "b_dim_1": 3}      # mutated like the Rust side
    assert dims["b"] == ["b_dim_0", "b_dim_1"] and dims["mu"] == []
    with pytest.raises(RuntimeError, match="inconsistent size"):
        _lib.PyVariable.new_variables(["a"], ["float64"], [[7]], {"county": 85}, {"a": ["county"]})
    with pytest.raises(RuntimeError, match="Unsupported item type"):
        _lib.PyVariable.new_variables(["a"], ["complex128"], [[1]], {}, {})
    with pytest.raises(RuntimeError, match="size unknown"):
        _lib.PyVariable.new_variables(["a"], ["float64"], [None], {}, {"a": ["nope"]})
    with pytest.raises(RuntimeError, match="number of dims"):
        _lib.PyVariable.new_variables(["a"], ["float64"], [[2, 2]], {}, {"a": ["x"]})


def test_lib_namespace_has_what_the_reference_python_layer_touches():
    """python/nutpie/sample.py:472-478 and __init__.py reference these at import time;
    compile_pymc.py:189-233 / compiled_pyfunc.py:72-105 construct the rest."""
    for name in ("PySampler", "PyMcModel", "LogpFunc", "ExpandFunc", "StanLibrary", "StanModel",
                 "PyNutsSettings", "PyMclmcSettings", "PyChainProgress", "ProgressType", "PyModel",
                 "PyVariable", "PyStorage", "PyTrace", "__version__", "store"):
        assert hasattr(_lib, name), name
    for name in ("LocalStore", "S3Store", "GCSStore", "AzureStore", "HTTPStore"):
        assert hasattr(_lib.store, name)
    for name in ("from_pymc", "from_stan", "from_pyfunc"):
        assert callable(getattr(_lib.PySampler, name))
    with pytest.raises(NotImplementedError):
        _lib.PySampler.from_stan()


def _numba_normal_cfuncs(mu, sigma):
    """A numba cfunc pair with the reference's exact signatures (compile_pymc.py:975-981,
    1018-1024): iid Normal(mu_i, sigma); `user_data` carries mu like the reference's record of
    shared-data pointers; expand appends the deterministic sum(x)."""
    numba = pytest.importorskip("numba")
    from numba import carray, cfunc, types

    n = len(mu)
    inv_var = 1.0 / sigma**2
    sig = types.int64(types.uint64, types.CPointer(types.double), types.CPointer(types.double),
                      types.CPointer(types.double), types.voidptr)

    @cfunc(sig, nopython=True)
    def logp(dim, x_, out_, logp_, ud_):
        if dim != n:
            return -1
        x = carray(x_, (n,))
        out = carray(out_, (n,))
        lp = carray(logp_, ())
        mu_ = carray(ud_, (n,), np.float64)
        acc = 0.0
        for i in range(n):
            r = x[i] - mu_[i]
            out[i] = -r * inv_var
            acc += r * r
        lp[()] = -0.5 * acc * inv_var
        if not np.isfinite(lp[()]):
            return 4
        return 0

    esig = types.int64(types.uint64, types.uint64, types.CPointer(types.double),
                       types.CPointer(types.double), types.voidptr)

    @cfunc(esig, nopython=True)
    def expand(dim, n_exp, x_, out_, ud_):
        if dim != n or n_exp != n + 1:
            return -1
        x = carray(x_, (n,))
        out = carray(out_, (n + 1,))
        s = 0.0
        for i in range(n):
            out[i] = x[i]
            s += x[i]
        out[n] = s
        return 0

    return logp, expand


def test_numba_cfunc_has_the_plugin_signature_and_runs_on_the_host():
    """The producer side of the boundary without PyMC (verified to compile here)."""
    mu = np.array([1.0, -2.0, 0.5])
    logp, expand = _numba_normal_cfuncs(mu, 2.0)
    fn = C.cast(logp.address, _lib.LOGP_FN)
    x = np.array([0.0, 0.0, 0.0])
    g = np.empty(3)
    lp = C.c_double()
    rc = fn(3, x.ctypes.data_as(C.POINTER(C.c_double)), g.ctypes.data_as(C.POINTER(C.c_double)),
            C.byref(lp), C.c_void_p(mu.ctypes.data))
    assert rc == 0
    np.testing.assert_allclose(g, mu / 4.0)
    np.testing.assert_allclose(lp.value, -0.5 * (mu**2).sum() / 4.0)
    assert fn(2, x.ctypes.data_as(C.POINTER(C.c_double)), g.ctypes.data_as(C.POINTER(C.c_double)),
              C.byref(lp), C.c_void_p(mu.ctypes.data)) == -1     # dim mismatch is fatal
    # the compiled-model wrapper builds the reference's objects call for call
    cm = NC.from_cfuncs(3, logp, expand, ["x", "total"], [(3,), ()], user_data=mu,
                        dims={"x": ("coord",)}, coords={"coord": ["a", "b", "c"]})
    m = cm._make_model(None)
    assert isinstance(m, _lib.PyMcModel) and m.dim == 3 and m.expand.expanded_dim == 4
    assert m.density.ptr == logp.address and m.density.user_data_ptr == mu.ctypes.data
    assert [(v.name, v.start_idx, v.end_idx) for v in m.variables] == [("x", 0, 3), ("total", 3, 4)]
    d, _ = m._descriptor()
    assert d.kind == _lib.MODEL_KINDS["host"] and d.host_logp == logp.address
    # host-side expand_vector through the C library (no GPU involved)
    out = m._expand(np.array([[1.0, 2.0, 3.0], [0.0, 0.0, 1.0]]))
    np.testing.assert_array_equal(out["total"], [6.0, 1.0])
    assert out["x"].shape == (2, 3)


def test_from_pyfunc_mirrors_compiled_pyfunc_py():
    def make_logp():
        return lambda x, scale: (-0.5 * float(x @ x) / scale, -x / scale)

    def make_expand(s1, s2, chain):
        return lambda x, scale: {"x": x, "r2": np.array(float(x @ x))}

    cm = nutpie_b200.from_pyfunc(4, make_logp, make_expand, [np.float64, np.float64], [(4,), ()],
                                 ["x", "r2"], shared_data={"scale": 2.0})
    assert cm.n_dim == 4 and cm.shapes == {"x": (4,), "r2": ()}
    with pytest.raises(ValueError, match="Unknown data variable"):
        cm.with_data(nope=1)
    m = cm.with_data(scale=4.0)._make_model(None)
    assert isinstance(m, _lib.PyModel) and m.dim == 4 and m.host_threads == 1
    # the C trampoline implements PyDensity::logp (src/pyfunc.rs:206-230)
    cb = m._trampoline()
    x, g, lp = np.ones(4), np.empty(4), C.c_double()
    ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert cb(4, ptr(x), ptr(g), C.byref(lp), None) == 0
    assert lp.value == -0.5 and np.array_equal(g, -x / 4.0)
    out = m._expand(np.arange(8.0).reshape(2, 4))
    assert out["r2"].tolist() == [14.0, 126.0]

    class Soft(Exception):
        is_recoverable = True

    def boom(kind):
        def f(x, **kw):
            if kind == "soft":
                raise Soft()
            if kind == "nan":
                return float("nan"), x
            if kind == "type":
                return "nope"
            raise KeyError("hard")
        return f

    for kind, rc in (("soft", 1), ("nan", 4), ("type", -2), ("hard", -1)):
        mm = nutpie_b200.from_pyfunc(4, lambda k=kind: boom(k), make_expand, [np.float64],
                                     [(4,)], ["x"])._make_model(None)
        assert mm._trampoline()(4, ptr(x), ptr(g), C.byref(lp), None) == rc, kind
        assert (mm.last_error is None) == (rc > 0)


def test_front_end_entry_points_say_what_to_use():
    with pytest.raises(ImportError, match="from_cfuncs"):
        nutpie_b200.compile_pymc_model(object())
    with pytest.raises(NotImplementedError):
        nutpie_b200.compile_stan_model(code="")


# ----------------------------------------------------------------------------- GPU
def _radon_pointer_model(d):
    """oracle_logp_radon / oracle_expand_radon (oracle/models.c) are host functions in the
    plug-in ABI: here they play the part of an existing compiled PyMC model."""
    J = d["n_county"]
    om = O.Model("radon", 2 * J + 5, y=d["y"], county=d["county"], floor=d["floor"], n_county=J)
    L = O.lib()
    names = ["intercept", "county_raw", "county_sd", "floor_effect", "county_floor_raw",
             "county_floor_sd", "sigma", "county_effect", "county_floor_effect"]
    shapes = [(), (J,), (), (), (J,), (), (), (J,), (J,)]

    class _UD:  # numpy-like holder so that user_data.ctypes.data is the RadonData pointer
        class ctypes:
            data = om.ud_ptr.value

    cm = NC.from_cfuncs(2 * J + 5, L.oracle_logp_radon, L.oracle_expand_radon, names, shapes,
                        user_data=_UD, dims={"county_raw": ("county",)},
                        coords={"county": np.arange(J)},
                        reparameterized_names=["county_sd", "county_floor_sd", "sigma"])
    return cm, om


@pytest.mark.gpu
def test_reference_abi_pointer_samples_like_the_device_density(radon_data):
    """The SAME radon density as a host pointer (NB200_MODEL_HOST) and as the hand-written CUDA
    density: same seed, same initial points => the same random streams; the trajectories agree
    to rounding for the first draws and statistically over the run."""
    d = radon_data
    J = d["n_county"]
    cm_host, om = _radon_pointer_model(d)
    cm_dev = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], J)
    n_chains, tune, draws = 32, 150, 150
    q0 = np.random.default_rng(5).uniform(-1, 1, (n_chains, cm_dev.n_dim))
    kw = dict(chains=n_chains, tune=tune, draws=draws, seed=11, q0=q0, progress_bar=False)
    a = nutpie_b200.sample(cm_host, **kw)
    b = nutpie_b200.sample(cm_dev, **kw)
    assert set(a.posterior) == set(b.posterior)
    for name in ("intercept", "sigma", "county_effect"):
        np.testing.assert_allclose(a.warmup_posterior[name][:, :3], b.warmup_posterior[name][:, :3],
                                   rtol=1e-7, atol=1e-9)
    sa, sb = a.sample_stats, b.sample_stats
    tot_a = sa["n_steps"].sum() + a.warmup_sample_stats["n_steps"].sum()
    tot_b = sb["n_steps"].sum() + b.warmup_sample_stats["n_steps"].sum()
    assert abs(tot_a / tot_b - 1) < 0.1
    assert abs(sa["step_size"][:, -1].mean() / sb["step_size"][:, -1].mean() - 1) < 0.1
    for name in ("intercept", "sigma"):
        xa, xb = a.posterior[name], b.posterior[name]
        se = np.hypot(xa.std() / np.sqrt(200), xb.std() / np.sqrt(200))  # >= 200 effective draws
        assert abs(xa.mean() - xb.mean()) < 5 * se, name
    assert a.dims["county_raw"] == ["county"] and a.posterior["county_effect"].shape == (n_chains, draws, J)


@pytest.mark.gpu
def test_numba_cfunc_samples_through_the_public_api():
    """tests/test_pymc.py:397-416 (`test_pymc_model_shared`): posterior means of N(mu, sigma)
    within tolerance, with the density compiled by numba exactly as compile_pymc_model does."""
    mu = np.array([-0.1, 10.0, 3.0])
    logp, expand = _numba_normal_cfuncs(mu, 3.0)
    cm = NC.from_cfuncs(3, logp, expand, ["x", "total"], [(3,), ()], user_data=mu)
    tr = nutpie_b200.sample(cm, chains=8, tune=300, draws=500, seed=3, progress_bar=False)
    x = tr.posterior["x"]
    assert x.shape == (8, 500, 3)
    np.testing.assert_allclose(x.mean(axis=(0, 1)), mu, atol=0.5)
    np.testing.assert_allclose(x.std(axis=(0, 1)), 3.0, rtol=0.15)
    np.testing.assert_allclose(tr.posterior["total"], x.sum(axis=-1))
    # new data through the same compiled functions (compile_pymc.py:136-161 `with_data`)
    mu2 = mu + 100.0
    tr2 = nutpie_b200.sample(cm.with_user_data(mu2), chains=4, tune=300, draws=300, seed=3,
                             progress_bar=False)
    np.testing.assert_allclose(tr2.posterior["x"].mean(axis=(0, 1)), mu2, atol=0.7)


@pytest.mark.gpu
def test_pyfunc_model_samples_and_reports_errors():
    """from_pyfunc (compiled_pyfunc.py:108-155) end to end; a density that raises is fatal and
    surfaces from wait() as RuntimeError with the partial trace kept (src/pyfunc.rs:100-116,
    src/wrapper.rs:1131-1136); `is_recoverable` errors are divergences."""
    def make_logp():
        return lambda x: (-0.5 * float((x - 2.0) @ (x - 2.0)), -(x - 2.0))

    def make_expand(s1, s2, chain):
        return lambda x: {"x": x}

    cm = nutpie_b200.from_pyfunc(2, make_logp, make_expand, [np.float64], [(2,)], ["x"],
                                 make_initial_point_fn=lambda seed: np.zeros(2))
    tr = nutpie_b200.sample(cm, chains=4, tune=100, draws=150, seed=1, progress_bar=False)
    x = tr.posterior["x"]
    assert x.shape == (4, 150, 2) and abs(x.mean() - 2.0) < 0.3 and abs(x.std() - 1.0) < 0.2

    calls = {"n": 0}

    def make_bad():
        def f(x):
            calls["n"] += 1
            if calls["n"] > 200:
                raise KeyError("the density broke")
            return -0.5 * float(x @ x), -x
        return f

    bad = nutpie_b200.from_pyfunc(2, make_bad, make_expand, [np.float64], [(2,)], ["x"])
    with pytest.raises(RuntimeError, match="error code") as ei:
        nutpie_b200.sample(bad, chains=2, tune=200, draws=200, seed=1, progress_bar=False)
    assert isinstance(ei.value.__cause__, KeyError)

    class Soft(Exception):
        is_recoverable = True

    def make_soft():
        def f(x):
            if x[0] > 1.5:
                raise Soft()
            return -0.5 * float(x @ x), -x
        return f

    soft = nutpie_b200.from_pyfunc(2, make_soft, make_expand, [np.float64], [(2,)], ["x"],
                                   make_initial_point_fn=lambda seed: np.zeros(2))
    tr = nutpie_b200.sample(soft, chains=2, tune=100, draws=100, seed=2, progress_bar=False)
    assert tr.posterior["x"][..., 0].max() <= 1.5
    assert tr.sample_stats["diverging"].sum() + tr.warmup_sample_stats["diverging"].sum() > 0


@pytest.mark.gpu
def test_fatal_return_code_stops_the_sampler_and_keeps_the_partial_trace(radon_data):
    """rc < 0 is not recoverable (src/pymc.rs:166-181): wait() raises, the rows finished so far
    stay readable — what `abort()` / `inspect()` return in the reference."""
    numba = pytest.importorskip("numba")
    from numba import carray, cfunc, types

    sig = types.int64(types.uint64, types.CPointer(types.double), types.CPointer(types.double),
                      types.CPointer(types.double), types.voidptr)

    calls = np.zeros(1, dtype=np.int64)  # user_data: evaluation counter (fails after 4000)

    @cfunc(sig, nopython=True)
    def logp(dim, x_, out_, logp_, ud_):
        x = carray(x_, (2,))
        out = carray(out_, (2,))
        lp = carray(logp_, ())
        n = carray(ud_, (1,), np.int64)
        n[0] += 1
        if n[0] > 4000:
            return -7
        out[0] = -x[0]
        out[1] = -x[1]
        lp[()] = -0.5 * (x[0] * x[0] + x[1] * x[1])
        return 0

    s = _lib.PyNutsSettings.Diag(4)
    s.update({"num_tune": 200, "num_draws": 2000})
    model = _lib.PyMcModel(_lib.LogpFunc(logp.address, calls.ctypes.data, (logp, calls)),
                           _lib.ExpandFunc(2, 2, 0, 0, None),
                           _lib.PyVariable.new_variables(["x"], ["float64"], [[2]], {}, {}),
                           2, {}, {}, lambda seed: np.zeros(2), None)
    smp = _lib.PySampler.from_pymc(s, 2, model, _lib.ProgressType.none(), None, 500,
                                   _lib.PyStorage.arrow(), n_chains=8)
    try:
        with pytest.raises(RuntimeError, match="Logp function returned error code: -7"):
            smp.wait()
        assert smp.is_finished()
        tr = smp.inspect()
        assert 0 < tr.rows_filled.max() < 2200
        n = int(tr.rows_filled.min())
        assert np.isfinite(tr.draws[:, :n]).all()
    finally:
        smp.close()
