"""CPU tests: the oracle against known answers and the reference's in-tree pins
(SURVEY.md §8c).  Bit-level parity with nuts-rs is unpinned; these are the pins
that exist."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as O


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_welford_matches_numpy():
    L = O.lib()
    rng = np.random.default_rng(0)
    x = rng.normal(size=(57, 6)) * np.array([1, 10, 0.1, 3, 1e3, 1e-3])
    mean, m2 = np.zeros(6), np.zeros(6)
    count = C.c_uint64(0)
    for row in x:
        L.oracle_welford_add(C.c_size_t(6), _p(mean), _p(m2), C.byref(count), _p(np.ascontiguousarray(row)))
    assert count.value == 57
    np.testing.assert_allclose(mean, x.mean(0), rtol=1e-12)
    np.testing.assert_allclose(m2 / 56, x.var(0, ddof=1), rtol=1e-11)


def test_mass_matrix_pins_normalizing_flow():
    """python/nutpie/normalizing_flow.py:1905-1915: one draw -> scale 1/sqrt(|g|)
    (variance 1/|g|); n draws -> scale sqrt(std q / std g) (variance std q / std g)."""
    L = O.lib()
    g = np.array([0.5, -4.0, 1e-3, 25.0])
    var = np.zeros(4)
    L.oracle_mass_matrix_init(C.c_size_t(4), _p(g), _p(var))
    diag = 1.0 / np.sqrt(np.abs(g))  # reference's one-draw `diag`
    np.testing.assert_allclose(var, diag ** 2, rtol=1e-15)
    rng = np.random.default_rng(1)
    q = rng.normal(size=(40, 4)) * np.array([1.0, 5.0, 0.2, 30.0])
    gr = rng.normal(size=(40, 4)) * np.array([2.0, 0.1, 9.0, 1.0])
    m2q = ((q - q.mean(0)) ** 2).sum(0)
    m2g = ((gr - gr.mean(0)) ** 2).sum(0)
    L.oracle_mass_matrix_update(C.c_size_t(4), C.c_int(1), _p(m2q), _p(m2g), C.c_uint64(40), _p(var))
    diag = np.sqrt(q.std(0) / gr.std(0))  # reference's n-draw `diag`
    np.testing.assert_allclose(var, diag ** 2, rtol=1e-13)
    # draw-based estimate ("draw_diag"): plain variance of the draws
    L.oracle_mass_matrix_update(C.c_size_t(4), C.c_int(0), _p(m2q), _p(m2g), C.c_uint64(40), _p(var))
    np.testing.assert_allclose(var, m2q / 40, rtol=1e-15)


def test_dual_averaging_matches_transcription():
    """Hoffman & Gelman dual averaging: k=.75, t0=10, gamma=.05,
    mu = log(10 eps0)."""
    L = O.lib()
    st = np.zeros(5)
    L.oracle_dual_average_init(_p(st), C.c_double(0.25))
    rng = np.random.default_rng(2)
    acc = rng.uniform(0.3, 1.0, size=50)
    log_step = log_bar = np.log(0.25)
    hbar, mu = 0.0, np.log(10 * 0.25)
    for n, a in enumerate(acc, start=1):
        L.oracle_dual_average_advance(_p(st), C.c_double(a), C.c_double(0.8), C.c_double(0.75),
                                      C.c_double(10.0), C.c_double(0.05))
        w = 1.0 / (n + 10.0)
        hbar = (1 - w) * hbar + w * (0.8 - a)
        log_step = mu - hbar * np.sqrt(n) / 0.05
        m = n ** -0.75
        log_bar = m * log_step + (1 - m) * log_bar
        np.testing.assert_allclose(st[:3], [log_step, log_bar, hbar], rtol=1e-12, atol=1e-14)
    assert st[4] == 51


def _leapfrog(model, q, p, g, var, psum, eps, direction, idx):
    L = O.lib()
    D = len(q)
    out = [np.zeros(D) for _ in range(4)]
    lp, kin = C.c_double(), C.c_double()
    rc = L.oracle_leapfrog(model.fn_ptr, model.ud_ptr, C.c_size_t(D), _p(q), _p(p), _p(g), _p(var),
                           _p(psum), C.c_double(eps), C.c_int(direction), C.c_int64(idx),
                           _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]), C.byref(lp), C.byref(kin))
    return rc, out, lp.value, kin.value


def test_leapfrog_closed_form_harmonic():
    """logp = -q^2/2, M = I: one leapfrog is the closed-form velocity-Verlet map."""
    m = O.Model("normal", 3)
    q = np.array([1.0, -2.0, 0.5]); p = np.array([0.3, 0.1, -1.0]); var = np.ones(3)
    g = -q
    eps = 0.1
    rc, (qn, pn, gn, sn), lp, kin = _leapfrog(m, q, p, g, var, p.copy(), eps, 1, 0)
    ph = p - 0.5 * eps * q
    q1 = q + eps * ph
    p1 = ph - 0.5 * eps * q1
    assert rc == 0
    np.testing.assert_allclose(qn, q1, rtol=1e-15)
    np.testing.assert_allclose(pn, p1, rtol=1e-15)
    np.testing.assert_allclose(gn, -q1, rtol=1e-15)
    np.testing.assert_allclose(sn, p + p1, rtol=1e-15)
    np.testing.assert_allclose(lp, -0.5 * (q1 ** 2).sum(), rtol=1e-15)
    np.testing.assert_allclose(kin, 0.5 * (p1 ** 2).sum(), rtol=1e-15)
    # backward step from idx 0 restarts the prefix sum (idx' == -1)
    rc, (qb, pb, gb, sb), _, _ = _leapfrog(m, q, p, g, var, p.copy(), eps, -1, 0)
    np.testing.assert_allclose(sb, pb, rtol=1e-15)
    # time reversal: forward then backward returns to the start
    rc, (q2, p2, _, _), _, _ = _leapfrog(m, qn, pn, gn, var, sn, eps, -1, 1)
    np.testing.assert_allclose(q2, q, atol=1e-15)
    np.testing.assert_allclose(p2, p, atol=1e-15)


def test_leapfrog_energy_error_is_second_order():
    m = O.Model("normal", 4, mu=1.0, sigma=2.0)
    rng = np.random.default_rng(3)
    q = rng.normal(size=4); p = rng.normal(size=4); var = np.array([0.5, 2.0, 1.0, 4.0])
    lp0, g0, _ = m.logp_grad(q)
    e0 = 0.5 * (p * var * p).sum() - lp0
    errs = []
    for eps in (0.2, 0.1, 0.05):
        rc, _, lp, kin = _leapfrog(m, q, p, g0, var, p.copy(), eps, 1, 0)
        errs.append(abs(kin - lp - e0))
    assert errs[0] / errs[1] > 3.0 and errs[1] / errs[2] > 3.0  # local error O(eps^3) .. O(eps^2)


def test_uturn_harmonic_oscillator_half_period():
    """1-D harmonic oscillator: the trajectory turns after half a period (pi)."""
    L = O.lib()
    m = O.Model("normal", 1)
    var = np.ones(1)
    eps = 0.05
    q, p = np.array([0.0]), np.array([1.0])
    g = -q
    psum = p.copy()
    p0, ps0 = p.copy(), psum.copy()
    turned_at = None
    idx = 0
    for n in range(1, 200):
        rc, (q, p, g, psum), _, _ = _leapfrog(m, q, p, g, var, psum, eps, 1, idx)
        idx += 1
        t = L.oracle_is_turning(C.c_size_t(1), C.c_int64(0), _p(p0), _p(ps0), C.c_int64(idx), _p(p),
                                _p(psum), _p(var))
        if t:
            turned_at = n * eps
            break
    assert turned_at is not None
    # rho = sum p ~ integral of cos -> sin(t)/eps, v_end = cos(t): turns when cos(t) < 0
    assert abs(turned_at - np.pi / 2) < 3 * eps


def test_uturn_cross_origin_case():
    """Span crossing the origin uses rho = psum_end + psum_start (Appendix A.3)."""
    L = O.lib()
    var = np.ones(2)
    p_l, s_l = np.array([1.0, 0.0]), np.array([2.0, 0.0])    # idx -2: sum over idx -1..-2
    p_r, s_r = np.array([1.0, 0.0]), np.array([3.0, 0.0])    # idx +2: sum over idx 0..2
    assert L.oracle_is_turning(C.c_size_t(2), C.c_int64(-2), _p(p_l), _p(s_l), C.c_int64(2), _p(p_r), _p(s_r), _p(var)) == 0
    p_r2 = np.array([-1.0, 0.0])
    assert L.oracle_is_turning(C.c_size_t(2), C.c_int64(-2), _p(p_l), _p(s_l), C.c_int64(2), _p(p_r2), _p(s_r), _p(var)) == 1


def test_radon_gradient_finite_differences(radon_data):
    d = radon_data
    J = d["n_county"]; D = 2 * J + 5
    m = O.Model("radon", D, y=d["y"], county=d["county"], floor=d["floor"], n_county=J)
    rng = np.random.default_rng(4)
    q = rng.normal(size=D) * 0.3
    lp, g, rc = m.logp_grad(q)
    assert rc == 0
    num = np.zeros(D)
    for i in range(D):
        e = np.zeros(D); e[i] = 1e-6
        num[i] = (m.logp_grad(q + e)[0] - m.logp_grad(q - e)[0]) / 2e-6
    np.testing.assert_allclose(g, num, rtol=2e-5, atol=2e-5)


def test_radon_logp_matches_scipy(radon_data):
    """The density restates PyMC's model.logp() including constants and Jacobians."""
    from scipy import stats

    d = radon_data
    J = d["n_county"]; D = 2 * J + 5
    m = O.Model("radon", D, y=d["y"], county=d["county"], floor=d["floor"], n_county=J)
    rng = np.random.default_rng(5)
    q = rng.normal(size=D) * 0.4
    lp, _, _ = m.logp_grad(q)
    ic, ra, lsa, fe, rb, lsb, lsg = q[0], q[1:J + 1], q[J + 1], q[J + 2], q[J + 3:2 * J + 3], q[2 * J + 3], q[2 * J + 4]
    sa, sb, sg = np.exp(lsa), np.exp(lsb), np.exp(lsg)
    mu = ic + (ra * sa)[d["county"]] + d["floor"] * (fe + (rb * sb)[d["county"]])
    ref = stats.norm(mu, sg).logpdf(d["y"]).sum()
    ref += stats.norm(0, 10).logpdf(ic) + stats.norm(0, 2).logpdf(fe)
    ref += stats.norm(0, 1).logpdf(ra).sum() + stats.norm(0, 1).logpdf(rb).sum()
    ref += stats.halfnorm(scale=1).logpdf(sa) + lsa + stats.halfnorm(scale=1).logpdf(sb) + lsb
    ref += stats.halfnorm(scale=1.5).logpdf(sg) + lsg
    np.testing.assert_allclose(lp, ref, rtol=1e-12)


def test_nonfinite_logp_return_codes():
    """compile_pymc.py:996-999: 4 = non-finite logp, 3 = non-finite gradient."""
    m = O.Model("funnel", 3)
    lp, g, rc = m.logp_grad(np.array([-800.0, 1.0, 1.0]))  # exp(1600) overflows
    assert rc in (3, 4)
    lp, g, rc = m.logp_grad(np.array([0.1, 1.0, 1.0]))
    assert rc == 0


def test_adam_step_size_rule_known_answer():
    """Adam on the log step size (nuts-rs stepsize/adam.rs as recalled; settings surfaced at
    src/wrapper.rs:347-392): 20-line numpy transcription vs oracle_adam_advance."""
    import ctypes as C

    L = O.lib()
    L.oracle_adam_advance.restype = None
    state = (C.c_double * 5)(np.log(0.1), np.log(0.1), 0.0, 0.0, 0.0)
    log_step, m, v, b1, b2, eps, lr, target = np.log(0.1), 0.0, 0.0, 0.9, 0.999, 1e-8, 0.05, 0.8
    rng = np.random.default_rng(0)
    for t in range(1, 40):
        acc = float(rng.uniform(0.3, 1.0))
        L.oracle_adam_advance(state, C.c_double(acc), C.c_double(target), C.c_double(lr))
        g = acc - target
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        log_step += lr * (m / (1 - b1**t)) / (np.sqrt(v / (1 - b2**t)) + eps)
        assert abs(state[0] - log_step) < 1e-13 and state[1] == state[0] and state[4] == t
    # too-high acceptance -> larger steps
    s2 = (C.c_double * 5)(0.0, 0.0, 0.0, 0.0, 0.0)
    for _ in range(10):
        L.oracle_adam_advance(s2, C.c_double(0.99), C.c_double(0.8), C.c_double(0.05))
    assert s2[0] > 0.3
