/*
 * oracle/models.c — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Host densities in the reference's own plug-in ABI
 *   int logp(size_t dim, const double* x, double* grad, double* logp, const void* user_data)
 * (src/pymc.rs:23-29; producer python/nutpie/compile_pymc.py:970-1006;
 * return codes compile_pymc.py:996-999: 0 ok, 3 non-finite gradient,
 * 4 non-finite logp, -1 dimension mismatch).  They are the host twins of the
 * device densities in nutpie_b200/csrc/models.cuh and can be handed to the
 * real nutpie via _lib.LogpFunc(ptr, user_data_ptr, keep_alive)
 * (python/nutpie/compile_pymc.py:197-201) on a machine that has it.
 *
 * The radon density restates what PyMC builds for the model in
 * README.md:53-88 / notebooks/pytensor_logp.md:57-88 (plain-Normal raw
 * effects, D = 2J+5): model.logp() with all normalising constants and the
 * log-Jacobians of the three HalfNormal log-transforms
 * (python/nutpie/compile_pymc.py:740-755).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "oracle.h"

#define LOG_2PI 1.8378770664093454835606594728112
#define HALF_LOG_2_OVER_PI (-0.22579135264472743236309761494744) /* 0.5*log(2/pi) */

static int finish(size_t dim, const double *grad, double logp, double *logp_out) {
    *logp_out = logp;
    if (!isfinite(logp)) return 4;
    for (size_t i = 0; i < dim; ++i)
        if (!isfinite(grad[i])) return 3;
    return 0;
}

/* logp = -1/2 sum ((x-mu)/sigma)^2 — Stan's `x ~ normal(mu, 1)` up to a
 * constant (README.md:148-163, tests/test_stan.py:16-24), and config 4. */
int oracle_logp_normal(size_t dim, const double *x, double *grad, double *logp_out,
                       const void *user_data) {
    const oracle_normal_data *d = (const oracle_normal_data *)user_data;
    double inv_var = 1.0 / (d->sigma * d->sigma);
    double acc = 0.0;
    for (size_t i = 0; i < dim; ++i) {
        double r = x[i] - d->mu;
        grad[i] = -r * inv_var;
        acc += r * r;
    }
    return finish(dim, grad, -0.5 * acc * inv_var, logp_out);
}

/* pm.HalfNormal("a") (sigma = 1) on PyMC's log-transformed scale, x = log a — the model of the
 * reference's seeded golden file tests/reference/test_deterministic_sampling_numba.txt
 * (tests/test_pymc.py:533-541): logp(x) = log(2/sqrt(2 pi)) - exp(2x)/2 + x, per coordinate. */
int oracle_logp_halfnormal(size_t dim, const double *x, double *grad, double *logp_out,
                           const void *user_data) {
    (void)user_data;
    double logp = 0.0;
    for (size_t i = 0; i < dim; ++i) {
        double e2 = exp(2.0 * x[i]);
        grad[i] = 1.0 - e2;
        logp += HALF_LOG_2_OVER_PI - 0.5 * e2 + x[i];
    }
    return finish(dim, grad, logp, logp_out);
}

/* Neal's funnel (docs/sample-stats.qmd:19-21 with 9 parameters as in
 * BASELINE.json): x0 = log_sigma ~ N(0,1), x[k] ~ N(0, exp(x0)). */
int oracle_logp_funnel(size_t dim, const double *x, double *grad, double *logp_out,
                       const void *user_data) {
    (void)user_data;
    double v = x[0];
    double e = exp(-2.0 * v);
    double ss = 0.0;
    for (size_t i = 1; i < dim; ++i) {
        ss += x[i] * x[i];
        grad[i] = -x[i] * e;
    }
    double n = (double)(dim - 1);
    grad[0] = -v + ss * e - n;
    double logp = -0.5 * v * v - 0.5 * ss * e - n * v;
    return finish(dim, grad, logp, logp_out);
}

/* Bayesian logistic regression, beta ~ N(0, 1), y_n ~ Bernoulli(logit^-1(X_n . beta)) — the
 * host twin of the run-time compiled CUDA density in tests/custom_densities.py (the device
 * analogue of a `from_pyfunc` user model, python/nutpie/compiled_pyfunc.py:108-155).
 * user_data: doubles [N, D, X row-major N*D, y N]. */
int oracle_logp_logreg(size_t dim, const double *x, double *grad, double *logp_out,
                       const void *user_data) {
    const double *d = (const double *)user_data;
    const size_t N = (size_t)d[0], D = (size_t)d[1];
    if (dim != D) return -1;
    const double *X = d + 2, *y = d + 2 + N * D;
    double logp = 0.0;
    for (size_t i = 0; i < D; ++i) {
        grad[i] = -x[i];
        logp += -0.5 * x[i] * x[i];
    }
    for (size_t n = 0; n < N; ++n) {
        double eta = 0.0;
        for (size_t i = 0; i < D; ++i) eta += X[n * D + i] * x[i];
        const double r = y[n] - 1.0 / (1.0 + exp(-eta));
        logp += y[n] * eta - (eta > 0.0 ? eta + log1p(exp(-eta)) : log1p(exp(eta)));
        for (size_t i = 0; i < D; ++i) grad[i] += X[n * D + i] * r;
    }
    return finish(dim, grad, logp, logp_out);
}

/* Radon, parameter order = PyMC value-variable order:
 *   [0] intercept, [1..J] county_raw, [J+1] log county_sd, [J+2] floor_effect,
 *   [J+3..2J+2] county_floor_raw, [2J+3] log county_floor_sd, [2J+4] log sigma */
int oracle_logp_radon(size_t dim, const double *x, double *grad, double *logp_out,
                      const void *user_data) {
    const oracle_radon_data *d = (const oracle_radon_data *)user_data;
    const int J = d->n_county, N = d->n_obs;
    if (dim != (size_t)(2 * J + 5)) return -1;
    const double intercept = x[0];
    const double *raw_a = x + 1;
    const double log_sd_a = x[J + 1];
    const double floor_eff = x[J + 2];
    const double *raw_b = x + J + 3;
    const double log_sd_b = x[2 * J + 3];
    const double log_sigma = x[2 * J + 4];
    const double sd_a = exp(log_sd_a), sd_b = exp(log_sd_b), sigma = exp(log_sigma);
    const double inv_sigma = 1.0 / sigma;

    for (size_t i = 0; i < dim; ++i) grad[i] = 0.0;
    double *g_a = grad + 1, *g_b = grad + J + 3;

    /* likelihood: y_i ~ N(mu_i, sigma) */
    double ss = 0.0, sum_e = 0.0, sum_fe = 0.0;
    for (int i = 0; i < N; ++i) {
        int c = d->county[i];
        double f = (double)d->floor[i];
        double mu = intercept + raw_a[c] * sd_a + f * (floor_eff + raw_b[c] * sd_b);
        double r = (d->y[i] - mu) * inv_sigma;
        double e = r * inv_sigma; /* d logp / d mu_i */
        ss += r * r;
        sum_e += e;
        sum_fe += f * e;
        g_a[c] += e;     /* accumulates E_c */
        g_b[c] += f * e; /* accumulates F_c */
    }
    double logp = -0.5 * ss - N * log_sigma - 0.5 * N * LOG_2PI;

    /* chain rule through county_effect = raw*sd and the priors */
    double d_log_sd_a = 0.0, d_log_sd_b = 0.0, ss_a = 0.0, ss_b = 0.0;
    for (int c = 0; c < J; ++c) {
        double Ec = g_a[c], Fc = g_b[c];
        d_log_sd_a += raw_a[c] * sd_a * Ec;
        d_log_sd_b += raw_b[c] * sd_b * Fc;
        g_a[c] = sd_a * Ec - raw_a[c];
        g_b[c] = sd_b * Fc - raw_b[c];
        ss_a += raw_a[c] * raw_a[c];
        ss_b += raw_b[c] * raw_b[c];
    }
    /* raw effects ~ N(0,1) */
    logp += -0.5 * ss_a - 0.5 * J * LOG_2PI;
    logp += -0.5 * ss_b - 0.5 * J * LOG_2PI;
    /* intercept ~ N(0,10) */
    logp += -0.5 * intercept * intercept / 100.0 - log(10.0) - 0.5 * LOG_2PI;
    grad[0] = sum_e - intercept / 100.0;
    /* floor_effect ~ N(0,2) */
    logp += -0.5 * floor_eff * floor_eff / 4.0 - log(2.0) - 0.5 * LOG_2PI;
    grad[J + 2] = sum_fe - floor_eff / 4.0;
    /* county_sd, county_floor_sd ~ HalfNormal(1), log-transformed */
    logp += HALF_LOG_2_OVER_PI - 0.5 * sd_a * sd_a + log_sd_a;
    grad[J + 1] = d_log_sd_a - sd_a * sd_a + 1.0;
    logp += HALF_LOG_2_OVER_PI - 0.5 * sd_b * sd_b + log_sd_b;
    grad[2 * J + 3] = d_log_sd_b - sd_b * sd_b + 1.0;
    /* sigma ~ HalfNormal(1.5), log-transformed */
    logp += HALF_LOG_2_OVER_PI - log(1.5) - 0.5 * sigma * sigma / 2.25 + log_sigma;
    grad[2 * J + 4] = ss - N - sigma * sigma / 2.25 + 1.0;

    return finish(dim, grad, logp, logp_out);
}

/* CpuLogpFunc::expand_vector for radon (src/pymc.rs:217-286; producer
 * compile_pymc.py:816-861): value vars on the constrained scale followed by
 * the two Deterministics.  out has 4J+5 entries:
 *   intercept, county_raw[J], county_sd, floor_effect, county_floor_raw[J],
 *   county_floor_sd, sigma, county_effect[J], county_floor_effect[J] */
int oracle_expand_radon(size_t dim, size_t expanded_dim, const double *x, double *out,
                        const void *user_data) {
    const oracle_radon_data *d = (const oracle_radon_data *)user_data;
    const int J = d->n_county;
    if (dim != (size_t)(2 * J + 5) || expanded_dim != (size_t)(4 * J + 5)) return -1;
    for (size_t i = 0; i < dim; ++i) out[i] = x[i];
    double sd_a = exp(x[J + 1]), sd_b = exp(x[2 * J + 3]);
    out[J + 1] = sd_a;
    out[2 * J + 3] = sd_b;
    out[2 * J + 4] = exp(x[2 * J + 4]);
    for (int c = 0; c < J; ++c) {
        out[dim + c] = x[1 + c] * sd_a;
        out[dim + J + c] = x[J + 3 + c] * sd_b;
    }
    return 0;
}
