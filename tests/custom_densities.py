"""CUDA sources of run-time compiled densities used by the tests (NB200_MODEL_CUSTOM,
include/nutpie_b200.h) and the data that goes with them.  Host twins: oracle/models.c."""
import numpy as np

# iid Normal(data[0], 1/sqrt(data[1])) — same density as the built-in NormalModel
NORMAL = r"""
__device__ int nb200_user_logp(const nb200_group& grp, int dim, const double* q, double* grad,
                               double* logp_partial, const double* data) {
    double acc = 0.0;
    for (int i = grp.tid; i < dim; i += grp.nthreads) {
        const double r = q[i] - data[0];
        grad[i] = -r * data[1];
        acc += r * r;
    }
    *logp_partial = -0.5 * acc * data[1];
    return 0;
}
"""

# Neal's funnel written with the group sum (twin: oracle_logp_funnel)
FUNNEL = r"""
__device__ int nb200_user_logp(const nb200_group& grp, int dim, const double* q, double* grad,
                               double* logp_partial, const double*) {
    const double v = q[0], e = exp(-2.0 * v);
    double ss = 0.0;
    for (int i = 1 + grp.tid; i < dim; i += grp.nthreads) {
        ss += q[i] * q[i];
        grad[i] = -q[i] * e;
    }
    ss = grp.sum(ss);
    const double n = (double)(dim - 1);
    if (grp.tid == 0) {
        grad[0] = -v + ss * e - n;
        *logp_partial = -0.5 * v * v - 0.5 * ss * e - n * v;
    }
    return 0;
}
"""

# Bayesian logistic regression (twin: oracle_logp_logreg); data = [N, D, X (N x D), y (N)],
# scratch = N doubles for the residuals
LOGREG = r"""
__device__ int nb200_user_logp(const nb200_group& grp, int dim, const double* q, double* grad,
                               double* logp_partial, const double* data) {
    const int N = (int)data[0];
    const double* X = data + 2;
    const double* y = X + (size_t)N * dim;
    double* resid = grp.scratch;
    double lp = 0.0;
    for (int n = grp.tid; n < N; n += grp.nthreads) {
        double eta = 0.0;
        for (int i = 0; i < dim; ++i) eta += X[(size_t)n * dim + i] * q[i];
        resid[n] = y[n] - 1.0 / (1.0 + exp(-eta));
        lp += y[n] * eta - (eta > 0.0 ? eta + log1p(exp(-eta)) : log1p(exp(eta)));
    }
    grp.sync();
    for (int i = grp.tid; i < dim; i += grp.nthreads) {
        double g = -q[i];
        for (int n = 0; n < N; ++n) g += X[(size_t)n * dim + i] * resid[n];
        grad[i] = g;
        lp += -0.5 * q[i] * q[i];
    }
    *logp_partial = lp;
    return 0;
}
"""

# recoverable error code (src/pymc.rs:178): the half-line q[0] > 1 is forbidden
WALL = r"""
__device__ int nb200_user_logp(const nb200_group& grp, int dim, const double* q, double* grad,
                               double* logp_partial, const double*) {
    double acc = 0.0;
    for (int i = grp.tid; i < dim; i += grp.nthreads) {
        grad[i] = -q[i];
        acc += q[i] * q[i];
    }
    *logp_partial = -0.5 * acc;
    return q[0] > 1.0 ? 1 : 0;
}
"""


def logreg_data(n_obs=200, dim=12, seed=3):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n_obs, dim)) / np.sqrt(dim)
    X[:, 0] = 1.0
    beta = rng.normal(size=dim)
    y = (rng.random(n_obs) < 1.0 / (1.0 + np.exp(-X @ beta))).astype(np.float64)
    return np.concatenate([[float(n_obs), float(dim)], X.ravel(), y])

# pm.HalfNormal("a") on the log scale (twin: oracle_logp_halfnormal) — the model of the
# reference's golden file tests/reference/test_deterministic_sampling_numba.txt
HALFNORMAL = r"""
__device__ int nb200_user_logp(const nb200_group& grp, int dim, const double* q, double* grad,
                               double* logp_partial, const double*) {
    double acc = 0.0;
    for (int i = grp.tid; i < dim; i += grp.nthreads) {
        const double e2 = exp(2.0 * q[i]);
        grad[i] = 1.0 - e2;
        acc += -0.22579135264472743236 - 0.5 * e2 + q[i];
    }
    *logp_partial = acc;
    return 0;
}
"""
