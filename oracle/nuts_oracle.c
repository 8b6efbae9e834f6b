/*
 * oracle/nuts_oracle.c — TEST INFRASTRUCTURE (CPU oracle), not product code.
 * See oracle/oracle.h for the parity status ("bit-level parity unpinned").
 *
 * A plain-C restatement of the per-chain NUTS loop that nutpie obtains from
 * nuts-rs 0.18.3 through nuts_rs::Sampler::new (src/wrapper.rs:977-1085) with
 * DiagNutsSettings (src/wrapper.rs:525-533).  Structure follows the crate's
 * published algorithm as summarised in SURVEY.md Appendix A:
 *   A.2 leapfrog + diagonal Euclidean Hamiltonian
 *   A.3 U-turn criterion on momentum prefix sums
 *   A.4 recursive tree doubling, multinomial / biased-progressive selection
 *   A.5 dual averaging, Welford draw+gradient estimators, window schedule,
 *       initial mass matrix and initial step-size search
 * The recursion mirrors the crate's tree `extend` so that the iterative CUDA
 * implementation can be checked against an independently shaped program.
 *
 * Deviation that cannot be avoided: random numbers come from the counter
 * based stream in oracle/philox.h, not from rand's ChaCha8 (Appendix A.6).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#include "oracle.h"
#include "philox.h"

#define VAR_LOWER 1e-20
#define VAR_UPPER 1e20

/* ------------------------------------------------------------------ state */
typedef struct State {
    double *q, *p, *g, *v, *p_sum;
    int64_t idx;
    double U, K; /* potential (= -logp) and kinetic energy */
    double E0;   /* energy of the trajectory's initial point */
    int refs;
} State;

typedef struct Pool {
    size_t dim;
    State **free_list;
    int n_free, cap;
} Pool;

static State *state_new(Pool *pool) {
    if (pool->n_free > 0) {
        State *s = pool->free_list[--pool->n_free];
        s->refs = 1;
        return s;
    }
    State *s = (State *)calloc(1, sizeof(State));
    double *buf = (double *)calloc(5 * pool->dim + 1, sizeof(double));
    s->q = buf;
    s->p = buf + pool->dim;
    s->g = buf + 2 * pool->dim;
    s->v = buf + 3 * pool->dim;
    s->p_sum = buf + 4 * pool->dim;
    s->refs = 1;
    return s;
}
static State *state_ref(State *s) {
    s->refs++;
    return s;
}
static void state_release(Pool *pool, State *s) {
    if (--s->refs > 0) return;
    if (pool->n_free == pool->cap) {
        pool->cap = pool->cap ? 2 * pool->cap : 64;
        pool->free_list = (State **)realloc(pool->free_list, pool->cap * sizeof(State *));
    }
    pool->free_list[pool->n_free++] = s;
}
static void pool_destroy(Pool *pool) {
    for (int i = 0; i < pool->n_free; ++i) {
        free(pool->free_list[i]->q);
        free(pool->free_list[i]);
    }
    free(pool->free_list);
}
static double state_energy(const State *s) { return s->K + s->U; }
static double state_energy_error(const State *s) { return state_energy(s) - s->E0; }

/* ------------------------------------------------------------------ chain */
typedef struct DualAverage {
    double log_step, log_step_adapted, hbar, mu;
    uint64_t count;
} DualAverage;

typedef struct RunVar {
    double *mean, *m2;
    uint64_t count;
} RunVar;

typedef struct Chain {
    size_t dim;
    const nb200_settings *st;
    nb200_logp_fn logp;
    const void *ud;
    uint64_t seed;
    uint32_t chain_id; /* global */
    const double *z_tape; /* [n_total][dim] or NULL */
    Pool pool;
    /* hamiltonian */
    double step_size;
    double *var, *inv_std;
    /* collectors for the current draw */
    double acc_sum, acc_sym_sum;
    uint64_t acc_count;
    uint32_t n_merge;
    uint32_t draw_index; /* rng "draw" coordinate of the running transition */
    /* step-size strategy */
    DualAverage da;
    double last_mean, last_sym;
    uint64_t last_n_steps;
    /* mass-matrix strategy */
    RunVar fg_draw, fg_grad, bg_draw, bg_grad;
    int has_initial_mass_matrix;
    uint64_t last_update;
    uint64_t total_steps;
    int fatal;
    /* DivergenceInfo of the running transition (store_divergences, python/nutpie/sample.py:641-646) */
    int div_valid;
    double *div; /* [4][dim]: start location, end location, start momentum, start gradient */
    /* low-rank strategy (st->adaptation == 1; oracle/lowrank.c): metric + the window of draws
     * and gradients, a deque [foreground-only part | background part] split at lr_split */
    int lr;
    double *lr_stds, *lr_vals, *lr_vecs;
    size_t lr_k, lr_max_rank;
    double *lr_win_q, *lr_win_g;
    size_t lr_len, lr_split, lr_cap;
} Chain;

/* the metric a leapfrog / U-turn check runs under: diagonal (var) or low rank */
typedef struct Metric {
    const double *var;
    int lr;
    const double *stds, *vals, *vecs;
    size_t k;
} Metric;
static Metric chain_metric(const Chain *c) {
    Metric m = {c->var, c->lr, c->lr_stds, c->lr_vals, c->lr_vecs, c->lr_k};
    return m;
}

/* ------------------------------------------------- component: dual average */
static void da_new(DualAverage *da, double initial_step) {
    da->log_step = log(initial_step);
    da->log_step_adapted = log(initial_step);
    da->hbar = 0.0;
    da->mu = log(10.0 * initial_step);
    da->count = 1;
}
static void da_advance(DualAverage *da, double accept_stat, double target, double k, double t0,
                       double gamma) {
    double count = (double)da->count;
    double w = 1.0 / (count + t0);
    da->hbar = (1.0 - w) * da->hbar + w * (target - accept_stat);
    da->log_step = da->mu - da->hbar * sqrt(count) / gamma;
    double mk = pow(count, -k);
    da->log_step_adapted = mk * da->log_step + (1.0 - mk) * da->log_step_adapted;
    da->count += 1;
}
/* Adam on the log step size (nuts-rs stepsize/adam.rs [recalled]; src/wrapper.rs:347-392):
 * state reuses the DualAverage slots — hbar = first moment, mu = second moment, count = t */
static void adam_new(DualAverage *da, double initial_step) {
    da->log_step = log(initial_step);
    da->log_step_adapted = da->log_step;
    da->hbar = 0.0;
    da->mu = 0.0;
    da->count = 0;
}
static void adam_advance(DualAverage *da, double accept_stat, double target, double lr) {
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    double grad = accept_stat - target;
    da->count += 1;
    da->hbar = b1 * da->hbar + (1.0 - b1) * grad;
    da->mu = b2 * da->mu + (1.0 - b2) * grad * grad;
    double t = (double)da->count;
    double m_hat = da->hbar / (1.0 - pow(b1, t));
    double v_hat = da->mu / (1.0 - pow(b2, t));
    da->log_step += lr * m_hat / (sqrt(v_hat) + eps);
    da->log_step_adapted = da->log_step;
}
static void step_new(const nb200_settings *st, DualAverage *da, double initial_step) {
    if (st->step_size_method == 1) adam_new(da, initial_step);
    else da_new(da, initial_step);
}
static void step_advance(const nb200_settings *st, DualAverage *da, double accept_stat) {
    if (st->step_size_method == 1) adam_advance(da, accept_stat, st->target_accept, st->adam_learning_rate);
    else da_advance(da, accept_stat, st->target_accept, st->da_k, st->da_t0, st->da_gamma);
}
void oracle_adam_advance(double state[5], double accept_stat, double target, double lr) {
    DualAverage da = {state[0], state[1], state[2], state[3], (uint64_t)state[4]};
    adam_advance(&da, accept_stat, target, lr);
    state[0] = da.log_step; state[1] = da.log_step_adapted; state[2] = da.hbar;
    state[3] = da.mu; state[4] = (double)da.count;
}
void oracle_dual_average_init(double state[5], double initial_step) {
    DualAverage da;
    da_new(&da, initial_step);
    state[0] = da.log_step; state[1] = da.log_step_adapted; state[2] = da.hbar;
    state[3] = da.mu; state[4] = (double)da.count;
}
void oracle_dual_average_advance(double state[5], double accept_stat, double target, double k,
                                 double t0, double gamma) {
    DualAverage da = {state[0], state[1], state[2], state[3], (uint64_t)state[4]};
    da_advance(&da, accept_stat, target, k, t0, gamma);
    state[0] = da.log_step; state[1] = da.log_step_adapted; state[2] = da.hbar;
    state[3] = da.mu; state[4] = (double)da.count;
}

/* --------------------------------------------- component: running variance */
void oracle_welford_add(size_t dim, double *mean, double *m2, uint64_t *count, const double *x) {
    *count += 1;
    if (*count == 1) {
        memcpy(mean, x, dim * sizeof(double));
        return;
    }
    double inv = 1.0 / (double)(*count);
    for (size_t i = 0; i < dim; ++i) {
        double diff = x[i] - mean[i];
        mean[i] += diff * inv;
        m2[i] += diff * (x[i] - mean[i]);
    }
}
static void runvar_reset(RunVar *r, size_t dim) {
    memset(r->mean, 0, dim * sizeof(double));
    memset(r->m2, 0, dim * sizeof(double));
    r->count = 0;
}
static void runvar_alloc(RunVar *r, size_t dim) {
    r->mean = (double *)calloc(dim + 1, sizeof(double));
    r->m2 = (double *)calloc(dim + 1, sizeof(double));
    r->count = 0;
}
static void runvar_free(RunVar *r) {
    free(r->mean);
    free(r->m2);
}

/* ------------------------------------------------ component: mass matrix */
/* pin: python/nutpie/normalizing_flow.py:1910-1914 — diag = sqrt(std q / std g)
 * as a scale, i.e. variance (M^-1) = sqrt(Var q / Var g); the common factor
 * 1/(n-1) cancels so the raw sums of squares are used. */
void oracle_mass_matrix_update(size_t dim, int use_grad, const double *m2_draw,
                               const double *m2_grad, uint64_t count, double *var_out) {
    for (size_t i = 0; i < dim; ++i) {
        double val = use_grad ? sqrt(m2_draw[i] / m2_grad[i]) : m2_draw[i] / (double)count;
        if (!isfinite(val) || val == 0.0) continue; /* keep previous */
        if (val < VAR_LOWER) val = VAR_LOWER;
        if (val > VAR_UPPER) val = VAR_UPPER;
        var_out[i] = val;
    }
}
/* pin: python/nutpie/normalizing_flow.py:1905-1909 — one-draw scale
 * 1/sqrt(|g|), i.e. variance = 1/|g| */
void oracle_mass_matrix_init(size_t dim, const double *grad, double *var_out) {
    for (size_t i = 0; i < dim; ++i) {
        double a = fabs(grad[i]);
        if (a < VAR_LOWER) a = VAR_LOWER;
        if (a > VAR_UPPER) a = VAR_UPPER;
        double val = 1.0 / a;
        if (!isfinite(val)) val = 1.0;
        var_out[i] = val;
    }
}
static void chain_set_var(Chain *c) {
    for (size_t i = 0; i < c->dim; ++i) c->inv_std[i] = sqrt(1.0 / c->var[i]);
}

/* ------------------------------------------------------- component: rng */
void oracle_rng_words(uint64_t seed, uint32_t chain, uint32_t draw, uint32_t purpose,
                      uint32_t index, uint64_t out[2]) {
    rng_u64x2(seed, chain, draw, purpose, index, &out[0], &out[1]);
}
void oracle_rng_normals(uint64_t seed, uint32_t chain, uint32_t draw, uint32_t purpose,
                        size_t dim, double *z) {
    for (size_t j = 0; 2 * j < dim; ++j) {
        uint64_t a, b;
        double z0, z1;
        rng_u64x2(seed, chain, draw, purpose, (uint32_t)j, &a, &b);
        rng_normal_pair(a, b, &z0, &z1);
        z[2 * j] = z0;
        if (2 * j + 1 < dim) z[2 * j + 1] = z1;
    }
}

/* -------------------------------------------------- component: leapfrog */
/* returns 0 ok, 1 divergence (recoverable logp error or energy error), <0 fatal */
static int leapfrog_raw(nb200_logp_fn logp, const void *ud, size_t dim, const Metric *mt,
                        const State *start, State *out, double step_size, int dir,
                        double max_energy_error) {
    const double *var = mt->var;
    double eps = (double)dir * step_size;
    out->E0 = start->E0;
    if (mt->lr) {
        for (size_t i = 0; i < dim; ++i) out->p[i] = start->p[i] + 0.5 * eps * start->g[i];
        oracle_lowrank_velocity(dim, mt->stds, mt->k, mt->vals, mt->vecs, out->p, out->v);
        for (size_t i = 0; i < dim; ++i) out->q[i] = start->q[i] + eps * out->v[i];
    } else
    for (size_t i = 0; i < dim; ++i) {
        double ph = start->p[i] + 0.5 * eps * start->g[i]; /* first momentum half step */
        out->p[i] = ph;
        double vh = var[i] * ph;                           /* velocity */
        out->q[i] = start->q[i] + eps * vh;                /* position step */
    }
    double lp = NAN;
    int rc = logp(dim, out->q, out->g, &lp, ud);
    out->idx = start->idx + dir;
    if (rc < 0) return rc; /* fatal, src/pymc.rs:166-181 */
    if (rc > 0) {          /* recoverable -> divergence */
        out->U = -lp;
        out->K = NAN;
        return 1;
    }
    out->U = -lp;
    double kin = 0.0;
    if (mt->lr) {
        for (size_t i = 0; i < dim; ++i) out->p[i] = out->p[i] + 0.5 * eps * out->g[i];
        oracle_lowrank_velocity(dim, mt->stds, mt->k, mt->vals, mt->vecs, out->p, out->v);
        for (size_t i = 0; i < dim; ++i) kin += out->p[i] * out->v[i];
    } else
    for (size_t i = 0; i < dim; ++i) {
        double pn = out->p[i] + 0.5 * eps * out->g[i]; /* second momentum half step */
        out->p[i] = pn;
        double vn = var[i] * pn;
        out->v[i] = vn;
        kin += pn * vn;
    }
    out->K = 0.5 * kin;
    if (out->idx == -1) {
        memcpy(out->p_sum, out->p, dim * sizeof(double));
    } else {
        for (size_t i = 0; i < dim; ++i) out->p_sum[i] = start->p_sum[i] + out->p[i];
    }
    double de = state_energy_error(out);
    if (de > max_energy_error || !isfinite(de)) return 1;
    return 0;
}

int oracle_leapfrog(nb200_logp_fn logp, const void *ud, size_t dim, const double *q,
                    const double *p, const double *g, const double *var, const double *p_sum,
                    double eps, int dir, int64_t idx, double *q_out, double *p_out,
                    double *g_out, double *p_sum_out, double *logp_out, double *kinetic_out) {
    Pool pool = {dim, NULL, 0, 0};
    State *a = state_new(&pool), *b = state_new(&pool);
    memcpy(a->q, q, dim * 8); memcpy(a->p, p, dim * 8); memcpy(a->g, g, dim * 8);
    memcpy(a->p_sum, p_sum, dim * 8);
    a->idx = idx; a->E0 = 0.0; a->U = 0.0; a->K = 0.0;
    Metric mt = {var, 0, NULL, NULL, NULL, 0};
    int rc = leapfrog_raw(logp, ud, dim, &mt, a, b, eps, dir, INFINITY);
    memcpy(q_out, b->q, dim * 8); memcpy(p_out, b->p, dim * 8); memcpy(g_out, b->g, dim * 8);
    memcpy(p_sum_out, b->p_sum, dim * 8);
    *logp_out = -b->U; *kinetic_out = b->K;
    state_release(&pool, a); state_release(&pool, b);
    pool_destroy(&pool);
    return rc;
}

/* ---------------------------------------------------- component: U-turn */
static int is_turning_raw(size_t dim, int64_t idx1, const double *p1, const double *psum1,
                          int64_t idx2, const double *p2, const double *psum2,
                          const double *var) {
    /* order along the trajectory */
    const double *ps, *pss, *pe, *pse;
    int64_t a, b;
    if (idx1 < idx2) { a = idx1; ps = p1; pss = psum1; b = idx2; pe = p2; pse = psum2; }
    else             { a = idx2; ps = p2; pss = psum2; b = idx1; pe = p1; pse = psum1; }
    double t_end = 0.0, t_start = 0.0;
    for (size_t i = 0; i < dim; ++i) {
        double rho;
        if (a >= 0 && b >= 0)      rho = pse[i] - pss[i] + ps[i];
        else if (b >= 0 && a < 0)  rho = pse[i] + pss[i];
        else                       rho = pss[i] - pse[i] + pe[i];
        t_end += rho * (var[i] * pe[i]);
        t_start += rho * (var[i] * ps[i]);
    }
    return (t_end < 0.0) | (t_start < 0.0);
}
int oracle_is_turning(size_t dim, int64_t idx1, const double *p1, const double *psum1,
                      int64_t idx2, const double *p2, const double *psum2, const double *var) {
    return is_turning_raw(dim, idx1, p1, psum1, idx2, p2, psum2, var);
}
/* the same criterion on stored velocities v = M^-1 p (any metric) */
static int is_turning_v(size_t dim, const State *s1, const State *s2) {
    const State *s = s1->idx < s2->idx ? s1 : s2, *e = s1->idx < s2->idx ? s2 : s1;
    int64_t a = s->idx, b = e->idx;
    double t_end = 0.0, t_start = 0.0;
    for (size_t i = 0; i < dim; ++i) {
        double rho;
        if (a >= 0 && b >= 0)      rho = e->p_sum[i] - s->p_sum[i] + s->p[i];
        else if (b >= 0 && a < 0)  rho = e->p_sum[i] + s->p_sum[i];
        else                       rho = s->p_sum[i] - e->p_sum[i] + e->p[i];
        t_end += rho * e->v[i];
        t_start += rho * s->v[i];
    }
    return (t_end < 0.0) | (t_start < 0.0);
}
static int is_turning(const Chain *c, const State *s1, const State *s2) {
    if (c->lr) return is_turning_v(c->dim, s1, s2);
    return is_turning_raw(c->dim, s1->idx, s1->p, s1->p_sum, s2->idx, s2->p, s2->p_sum, c->var);
}

/* -------------------------------------------------------- trajectory init */
static void fill_momentum_normals(Chain *c, uint32_t purpose, uint32_t draw, double *z) {
    if (purpose == RNG_MOMENTUM && c->z_tape) {
        memcpy(z, c->z_tape + (size_t)draw * c->dim, c->dim * sizeof(double));
        return;
    }
    oracle_rng_normals(c->seed, c->chain_id, draw, purpose, c->dim, z);
}
/* fresh momentum p = inv_std * z, v = var*p, K, p_sum = p, idx = 0, E0 = E */
static void initialize_trajectory(Chain *c, State *s, uint32_t purpose, uint32_t draw) {
    double *z = s->p_sum; /* scratch, overwritten below */
    fill_momentum_normals(c, purpose, draw, z);
    double kin = 0.0;
    if (c->lr) {
        oracle_lowrank_momentum(c->dim, c->lr_stds, c->lr_k, c->lr_vals, c->lr_vecs, z, s->p);
        oracle_lowrank_velocity(c->dim, c->lr_stds, c->lr_k, c->lr_vals, c->lr_vecs, s->p, s->v);
        for (size_t i = 0; i < c->dim; ++i) kin += s->p[i] * s->v[i];
    } else
    for (size_t i = 0; i < c->dim; ++i) {
        double p = c->inv_std[i] * z[i];
        s->p[i] = p;
        double v = c->var[i] * p;
        s->v[i] = v;
        kin += p * v;
    }
    s->K = 0.5 * kin;
    s->idx = 0;
    s->E0 = state_energy(s);
    memcpy(s->p_sum, s->p, c->dim * sizeof(double));
}

/* leapfrog + acceptance-rate collector (nuts-rs AcceptanceRateCollector) */
static int chain_leapfrog(Chain *c, const State *start, State *out, int dir) {
    Metric mt = chain_metric(c);
    int rc = leapfrog_raw(c->logp, c->ud, c->dim, &mt, start, out, c->step_size, dir,
                          c->st->max_energy_error);
    if (rc < 0) return rc;
    c->acc_count += 1;
    c->total_steps += 1;
    if (rc == 1 && c->div) {
        c->div_valid = 1;
        memcpy(c->div, start->q, c->dim * 8);
        memcpy(c->div + c->dim, out->q, c->dim * 8);
        memcpy(c->div + 2 * c->dim, start->p, c->dim * 8);
        memcpy(c->div + 3 * c->dim, start->g, c->dim * 8);
    }
    if (rc == 0) {
        double de = state_energy_error(out);
        double w = exp(-de);
        double a = w < 1.0 ? w : 1.0;
        c->acc_sum += a;
        c->acc_sym_sum += 2.0 * a / (1.0 + w);
    }
    return rc;
}

/* ------------------------------------------------------------------- tree */
typedef struct Tree {
    State *left, *right, *draw;
    double log_size;
    uint32_t depth;
    int is_main;
} Tree;

enum { EXT_OK = 0, EXT_TURNING = 1, EXT_DIVERGING = 2, EXT_ERR = 3 };

static void tree_release(Chain *c, Tree *t) {
    state_release(&c->pool, t->left);
    state_release(&c->pool, t->right);
    state_release(&c->pool, t->draw);
}
static double logaddexp(double a, double b) {
    if (a == b) return a + M_LN2;
    double diff = a - b;
    if (diff > 0) return a + log1p(exp(-diff));
    if (diff < 0) return b + log1p(exp(diff));
    return diff; /* NaN */
}

/* merge `other` into `self` (consumes other); multinomial pick inside
 * sub-trees, biased progressive pick for the main tree (Appendix A.4) */
static void merge_into(Chain *c, Tree *self, Tree *other, int dir) {
    if (dir > 0) {
        state_release(&c->pool, self->right);
        self->right = state_ref(other->right);
    } else {
        state_release(&c->pool, self->left);
        self->left = state_ref(other->left);
    }
    double log_size = logaddexp(self->log_size, other->log_size);
    double self_log_size = self->is_main ? self->log_size : log_size;
    uint64_t a, b;
    rng_u64x2(c->seed, c->chain_id, c->draw_index, RNG_MERGE, c->n_merge, &a, &b);
    c->n_merge += 1;
    double u = rng_u01(a);
    if (other->log_size >= self_log_size || u < exp(other->log_size - self_log_size)) {
        state_release(&c->pool, self->draw);
        self->draw = state_ref(other->draw);
    }
    self->depth += 1;
    self->log_size = log_size;
    tree_release(c, other);
}

/* Extend `self` by a sub-tree of equal depth in direction dir.  On EXT_OK /
 * EXT_TURNING with merged=1 the new sub-tree has been merged into self. */
static int tree_extend(Chain *c, Tree *self, int dir, int check_turning, int *merged) {
    *merged = 0;
    /* single step from the end of self */
    const State *start = dir > 0 ? self->right : self->left;
    State *end = state_new(&c->pool);
    int rc = chain_leapfrog(c, start, end, dir);
    if (rc != 0) {
        state_release(&c->pool, end);
        return rc < 0 ? EXT_ERR : EXT_DIVERGING;
    }
    Tree other = {end, state_ref(end), state_ref(end), -state_energy_error(end), 0, 0};
    while (other.depth < self->depth) {
        int sub_merged;
        int r = tree_extend(c, &other, dir, check_turning, &sub_merged);
        if (r != EXT_OK) { /* turning inside the new sub-tree, divergence or error */
            tree_release(c, &other);
            return r;
        }
    }
    const State *first = dir > 0 ? self->left : other.left;
    const State *last = dir > 0 ? other.right : self->right;
    int turning = 0;
    if (check_turning) {
        turning = is_turning(c, first, last);
        if (self->depth > 0) {
            if (!turning) turning = is_turning(c, self->right, other.right);
            if (!turning) turning = is_turning(c, self->left, other.left);
        }
    }
    merge_into(c, self, &other, dir);
    *merged = 1;
    return turning ? EXT_TURNING : EXT_OK;
}

typedef struct SampleInfo {
    uint32_t depth;
    int diverging, reached_maxdepth;
} SampleInfo;

/* one NUTS transition; `init` holds (q, g, U); returns the selected state */
static State *nuts_draw(Chain *c, State *init, uint32_t draw, SampleInfo *info) {
    c->draw_index = draw;
    c->n_merge = 0;
    initialize_trajectory(c, init, RNG_MOMENTUM, draw);
    c->acc_sum = c->acc_sym_sum = 0.0;
    c->acc_count = 0;
    Tree tree = {state_ref(init), state_ref(init), state_ref(init), 0.0, 0, 1};
    info->diverging = 0;
    info->reached_maxdepth = 0;
    int done = 0;
    while (tree.depth < c->st->maxdepth && !done) {
        uint64_t a, b;
        rng_u64x2(c->seed, c->chain_id, draw, RNG_DIRECTION, tree.depth, &a, &b);
        int dir = (a & 1) ? 1 : -1;
        int check = c->st->check_turning && tree.depth >= c->st->mindepth;
        int merged;
        int r = tree_extend(c, &tree, dir, check, &merged);
        switch (r) {
        case EXT_OK: break;
        case EXT_TURNING: done = 1; break;
        case EXT_DIVERGING: info->diverging = 1; done = 1; break;
        default: c->fatal = 1; done = 1; break;
        }
    }
    if (!done) info->reached_maxdepth = 1;
    info->depth = tree.depth;
    State *out = state_ref(tree.draw);
    tree_release(c, &tree);
    return out;
}

/* ------------------------------------------------------ step-size search */
/* nuts-rs step-size Strategy::init: one trial leapfrog decides the search
 * direction, then double/halve until the acceptance statistic crosses the
 * target (Appendix A.5). */
static void step_size_init(Chain *c, const State *point, uint32_t rng_draw) {
    const nb200_settings *st = c->st;
    if (st->step_size_method == 2) {
        c->step_size = st->fixed_step_size;
        return;
    }
    State *s = state_new(&c->pool);
    memcpy(s->q, point->q, c->dim * 8);
    memcpy(s->g, point->g, c->dim * 8);
    s->U = point->U;
    initialize_trajectory(c, s, RNG_STEP_INIT, rng_draw);
    State *nxt = state_new(&c->pool);
    uint64_t keep_count = c->acc_count, keep_total = c->total_steps;
    double keep_sum = c->acc_sum, keep_sym = c->acc_sym_sum;

    c->step_size = st->initial_step;
    int found = 0;
    c->acc_sum = 0; c->acc_count = 0;
    int rc = chain_leapfrog(c, s, nxt, 1);
    if (rc == 0) {
        double accept = c->acc_sum;
        int dir = accept > st->target_accept ? 1 : -1;
        for (int it = 0; it < 100; ++it) {
            c->acc_sum = 0; c->acc_count = 0;
            rc = chain_leapfrog(c, s, nxt, dir);
            if (rc != 0) {
                c->step_size = st->initial_step;
                found = -1; /* give up, keep the existing dual average */
                break;
            }
            accept = c->acc_sum;
            if (dir > 0) {
                if (accept <= st->target_accept || c->step_size > 1e5) { found = 1; break; }
                c->step_size *= 2.0;
            } else {
                if (accept >= st->target_accept || c->step_size < 1e-10) { found = 1; break; }
                c->step_size /= 2.0;
            }
        }
        if (found == 0) {
            c->step_size = st->initial_step;
            found = 1;
        }
        if (found == 1) step_new(st, &c->da, c->step_size);
    }
    c->acc_count = keep_count; c->acc_sum = keep_sum; c->acc_sym_sum = keep_sym;
    c->total_steps = keep_total; /* search leapfrogs are not trajectory steps */
    state_release(&c->pool, s);
    state_release(&c->pool, nxt);
}

static double clamp_step(const Chain *c, double step) {
    double m = c->st->max_step_size;
    return (m > 0 && step > m) ? m : step;
}

/* ------------------------------------------------------------ adaptation */
static int update_mass_matrix(Chain *c) {
    if (c->fg_draw.count < 3) return 0;
    oracle_mass_matrix_update(c->dim, c->st->use_grad_based_estimate, c->fg_draw.m2,
                              c->fg_grad.m2, c->fg_draw.count, c->var);
    chain_set_var(c);
    return 1;
}

/* low-rank strategy: window deque + refresh of the metric (oracle/lowrank.c) */
static void lr_push(Chain *c, const double *q, const double *g) {
    if (c->lr_len == c->lr_cap) {
        c->lr_cap = c->lr_cap ? 2 * c->lr_cap : 64;
        c->lr_win_q = (double *)realloc(c->lr_win_q, c->lr_cap * c->dim * 8);
        c->lr_win_g = (double *)realloc(c->lr_win_g, c->lr_cap * c->dim * 8);
    }
    memcpy(c->lr_win_q + c->lr_len * c->dim, q, c->dim * 8);
    memcpy(c->lr_win_g + c->lr_len * c->dim, g, c->dim * 8);
    c->lr_len += 1;
}
static void lr_switch(Chain *c) { /* drop the foreground-only part; the rest becomes it */
    size_t keep = c->lr_len - c->lr_split;
    memmove(c->lr_win_q, c->lr_win_q + c->lr_split * c->dim, keep * c->dim * 8);
    memmove(c->lr_win_g, c->lr_win_g + c->lr_split * c->dim, keep * c->dim * 8);
    c->lr_len = keep;
    c->lr_split = keep;
}
static int lr_update(Chain *c) {
    if (c->lr_len < 3) return 0;
    size_t k = 0;
    if (oracle_lowrank_update(c->dim, c->lr_len, c->lr_win_q, c->lr_win_g, c->st->mass_matrix_gamma,
                              c->st->mass_matrix_eigval_cutoff, c->lr_max_rank, c->lr_stds,
                              c->lr_vals, c->lr_vecs, &k) != 0)
        return 0;
    c->lr_k = k;
    return 1;
}

/* nuts-rs GlobalStrategy::adapt, called after every draw (Appendix A.5) */
static void adapt(Chain *c, uint64_t t, const State *draw, const SampleInfo *info) {
    const nb200_settings *st = c->st;
    c->last_mean = c->acc_count ? c->acc_sum / (double)c->acc_count : 0.0;
    c->last_sym = c->acc_count ? c->acc_sym_sum / (double)c->acc_count : 0.0;
    c->last_n_steps = c->acc_count;
    uint64_t num_tune = st->num_tune;
    if (t >= num_tune) return;
    int fixed = st->step_size_method == 2;
    uint64_t early_end = (uint64_t)ceil(st->early_window * (double)num_tune);
    uint64_t sw = (uint64_t)ceil(st->step_size_window * (double)num_tune);
    uint64_t final_window = sw > num_tune ? 0 : num_tune - sw;

    if (t < final_window) {
        int is_early = t < early_end;
        uint64_t switch_freq =
            is_early ? st->early_mass_matrix_switch_freq : st->mass_matrix_switch_freq;
        /* DrawGradCollector: which draws feed the estimators */
        int is_good = info->diverging ? (llabs(draw->idx) > 4) : (draw->idx != 0);
        if (is_good && c->lr) {
            lr_push(c, draw->q, draw->g);
        } else if (is_good) {
            oracle_welford_add(c->dim, c->fg_draw.mean, c->fg_draw.m2, &c->fg_draw.count, draw->q);
            oracle_welford_add(c->dim, c->fg_grad.mean, c->fg_grad.m2, &c->fg_grad.count, draw->g);
            oracle_welford_add(c->dim, c->bg_draw.mean, c->bg_draw.m2, &c->bg_draw.count, draw->q);
            oracle_welford_add(c->dim, c->bg_grad.mean, c->bg_grad.m2, &c->bg_grad.count, draw->g);
        }
        int could_switch = (c->lr ? c->lr_len - c->lr_split : c->bg_draw.count) >= switch_freq;
        int is_late = switch_freq + t > final_window;
        int force_update = 0;
        if (could_switch && !is_late && c->lr) {
            lr_switch(c);
            force_update = 1;
        } else if (could_switch && !is_late) {
            RunVar td = c->fg_draw, tg = c->fg_grad;
            c->fg_draw = c->bg_draw; c->fg_grad = c->bg_grad;
            c->bg_draw = td; c->bg_grad = tg;
            runvar_reset(&c->bg_draw, c->dim);
            runvar_reset(&c->bg_grad, c->dim);
            force_update = 1;
        }
        int did_change = 0;
        if (force_update || (t - c->last_update >= st->mass_matrix_update_freq))
            did_change = c->lr ? lr_update(c) : update_mass_matrix(c);
        if (did_change) c->last_update = t;
        if (!fixed) step_advance(st, &c->da, is_late ? c->last_sym : c->last_mean);
        if (did_change && c->has_initial_mass_matrix) {
            c->has_initial_mass_matrix = 0;
            step_size_init(c, draw, (uint32_t)t);
        } else if (!fixed) {
            c->step_size = clamp_step(c, exp(c->da.log_step));
        }
        return;
    }
    if (fixed) return;
    step_advance(st, &c->da, c->last_sym);
    if (t == num_tune - 1)
        c->step_size = clamp_step(c, exp(c->da.log_step_adapted));
    else
        c->step_size = clamp_step(c, exp(c->da.log_step));
}

/* ------------------------------------------------------------- chain run */
static int chain_init_position(Chain *c, State *s, const double *q0, const double *init_mean) {
    const nb200_settings *st = c->st;
    int tries = q0 ? 1 : (st->num_try_init > 0 ? st->num_try_init : 1);
    for (int attempt = 0; attempt < tries; ++attempt) {
        if (q0) {
            memcpy(s->q, q0, c->dim * 8);
        } else {
            for (size_t j = 0; 2 * j < c->dim; ++j) {
                uint64_t a, b;
                rng_u64x2(c->seed, c->chain_id, (uint32_t)attempt, RNG_INIT_POS, (uint32_t)j, &a, &b);
                double e0, e1;
                if (st->init_kind == 1) {
                    rng_normal_pair(a, b, &e0, &e1);
                } else {
                    e0 = st->init_radius * (2.0 * rng_u01(a) - 1.0);
                    e1 = st->init_radius * (2.0 * rng_u01(b) - 1.0);
                }
                size_t i0 = 2 * j, i1 = 2 * j + 1;
                s->q[i0] = (init_mean ? init_mean[i0] : 0.0) + e0;
                if (i1 < c->dim) s->q[i1] = (init_mean ? init_mean[i1] : 0.0) + e1;
            }
        }
        double lp = NAN;
        int rc = c->logp(c->dim, s->q, s->g, &lp, c->ud);
        if (rc < 0) return NB200_ELOGP;
        if (rc == 0 && isfinite(lp)) {
            s->U = -lp;
            return 0;
        }
    }
    return NB200_EINIT;
}

static int run_chain(const nb200_settings *st, nb200_logp_fn logp, const void *ud, size_t dim,
                     uint32_t chain_id, const double *q0, const double *init_mean,
                     const double *z_tape, size_t n_rows, size_t sdim, double *draws,
                     double *stats, double *grads, double *mminv, double *divs, double *eigvals,
                     uint64_t *steps_out) {
    Chain c;
    memset(&c, 0, sizeof(c));
    c.dim = dim; c.st = st; c.logp = logp; c.ud = ud; c.seed = st->seed;
    c.chain_id = chain_id; c.z_tape = z_tape;
    c.pool.dim = dim;
    c.var = (double *)calloc(dim + 1, 8);
    c.inv_std = (double *)calloc(dim + 1, 8);
    c.div = divs ? (double *)calloc(4 * dim + 1, 8) : NULL;
    runvar_alloc(&c.fg_draw, dim); runvar_alloc(&c.fg_grad, dim);
    runvar_alloc(&c.bg_draw, dim); runvar_alloc(&c.bg_grad, dim);
    c.lr = st->adaptation == 1;
    c.lr_max_rank = st->mass_matrix_max_rank < dim ? st->mass_matrix_max_rank : dim;
    if (c.lr) {
        c.lr_stds = (double *)calloc(dim + 1, 8);
        c.lr_vals = (double *)calloc(c.lr_max_rank + 1, 8);
        c.lr_vecs = (double *)calloc(c.lr_max_rank * dim + 1, 8);
    }

    State *cur = state_new(&c.pool);
    int rc = chain_init_position(&c, cur, q0, init_mean);
    if (rc == 0) {
        /* GlobalStrategy::init: mass matrix from |grad|, estimators seeded with the
         * initial point, then the step-size search */
        oracle_mass_matrix_init(dim, cur->g, c.var);
        chain_set_var(&c);
        oracle_welford_add(dim, c.fg_draw.mean, c.fg_draw.m2, &c.fg_draw.count, cur->q);
        oracle_welford_add(dim, c.bg_draw.mean, c.bg_draw.m2, &c.bg_draw.count, cur->q);
        oracle_welford_add(dim, c.fg_grad.mean, c.fg_grad.m2, &c.fg_grad.count, cur->g);
        oracle_welford_add(dim, c.bg_grad.mean, c.bg_grad.m2, &c.bg_grad.count, cur->g);
        if (c.lr) { /* low rank [recalled]: identity metric, the window seeded with the initial point */
            for (size_t i = 0; i < dim; ++i) c.lr_stds[i] = 1.0;
            c.lr_k = 0;
            lr_push(&c, cur->q, cur->g);
        }
        c.has_initial_mass_matrix = 1;
        step_new(st, &c.da, st->initial_step);
        c.step_size = st->initial_step;
        step_size_init(&c, cur, 0xFFFFFFFFu);

        uint64_t n_total = st->num_tune + st->num_draws;
        for (uint64_t t = 0; t < n_total && !c.fatal; ++t) {
            SampleInfo info;
            /* step_size_jitter: factor uniform in [1 - j, 1 + j] on the step used for this draw */
            double step_base = c.step_size;
            if (st->step_size_jitter > 0.0) {
                uint64_t ja, jb;
                rng_u64x2(c.seed, c.chain_id, (uint32_t)t, RNG_JITTER, 0u, &ja, &jb);
                c.step_size = step_base * (1.0 + st->step_size_jitter * (2.0 * rng_u01(ja) - 1.0));
            }
            double step_used = c.step_size;
            c.div_valid = 0;
            int store = st->save_warmup || t >= st->num_tune;
            size_t row = st->save_warmup ? t : t - st->num_tune;
            /* store_mass_matrix: mass_matrix_inv (diag) | mass_matrix_stds + mass_matrix_eigvals
             * (low rank; eigvals NaN-padded to mass_matrix_max_rank) */
            if (store && mminv) memcpy(mminv + row * sdim, c.lr ? c.lr_stds : c.var, sdim * 8);
            if (store && eigvals && c.lr)
                for (size_t k = 0; k < st->mass_matrix_max_rank; ++k)
                    eigvals[row * st->mass_matrix_max_rank + k] = k < c.lr_k ? c.lr_vals[k] : NAN;
            State *nxt = nuts_draw(&c, cur, (uint32_t)t, &info);
            c.step_size = step_base;
            if (c.fatal) { state_release(&c.pool, nxt); break; }
            if (store && row < n_rows && divs) {
                double *o = divs + row * 4 * sdim;
                for (size_t k = 0; k < 4; ++k)
                    for (size_t i = 0; i < sdim; ++i)
                        o[k * sdim + i] = c.div_valid && info.diverging ? c.div[k * dim + i] : NAN;
            }
            adapt(&c, t, nxt, &info);
            if (store && row < n_rows) {
                memcpy(draws + row * sdim, nxt->q, sdim * 8);
                if (grads) memcpy(grads + row * sdim, nxt->g, sdim * 8);
                double *s = stats + row * NB200_NSTAT;
                s[NB200_STAT_DEPTH] = info.depth;
                s[NB200_STAT_MAXDEPTH_REACHED] = info.reached_maxdepth;
                s[NB200_STAT_INDEX_IN_TRAJECTORY] = (double)nxt->idx;
                s[NB200_STAT_LOGP] = -nxt->U;
                s[NB200_STAT_ENERGY] = state_energy(nxt);
                s[NB200_STAT_ENERGY_ERROR] = state_energy_error(nxt);
                s[NB200_STAT_DIVERGING] = info.diverging;
                s[NB200_STAT_STEP_SIZE] = step_used;
                s[NB200_STAT_STEP_SIZE_BAR] = exp(c.da.log_step_adapted);
                s[NB200_STAT_N_STEPS] = (double)c.last_n_steps;
                s[NB200_STAT_MEAN_TREE_ACCEPT] = c.last_mean;
                s[NB200_STAT_MEAN_TREE_ACCEPT_SYM] = c.last_sym;
                s[NB200_STAT_TUNING] = t < st->num_tune;
                s[NB200_STAT_DRAW] = (double)t;
                s[NB200_STAT_CHAIN] = (double)chain_id;
                s[NB200_STAT_RESERVED] = 0.0;
            }
            state_release(&c.pool, cur);
            cur = nxt;
        }
        if (c.fatal) rc = NB200_ELOGP;
    }
    *steps_out = c.total_steps;
    state_release(&c.pool, cur);
    pool_destroy(&c.pool);
    free(c.var); free(c.inv_std); free(c.div);
    free(c.lr_stds); free(c.lr_vals); free(c.lr_vecs); free(c.lr_win_q); free(c.lr_win_g);
    runvar_free(&c.fg_draw); runvar_free(&c.fg_grad);
    runvar_free(&c.bg_draw); runvar_free(&c.bg_grad);
    return rc;
}

typedef struct Job {
    const nb200_settings *st;
    nb200_logp_fn logp;
    const void *ud;
    uint64_t dim, n_chains, chain_id_offset;
    const double *q0, *init_mean, *z_tape;
    double *draws, *stats, *gradients, *mminv, *divs, *eigvals;
    size_t n_total, n_rows, sdim;
    atomic_long next;  /* dynamic schedule over chains */
    atomic_ullong steps;
    atomic_int err;
} Job;

static void *worker(void *arg) {
    Job *j = (Job *)arg;
    for (;;) {
        long ci = atomic_fetch_add(&j->next, 1);
        if (ci >= (long)j->n_chains) break;
        uint64_t s = 0;
        int rc = run_chain(j->st, j->logp, j->ud, j->dim, (uint32_t)(j->chain_id_offset + ci),
                           j->q0 ? j->q0 + ci * j->dim : NULL, j->init_mean,
                           j->z_tape ? j->z_tape + (size_t)ci * j->n_total * j->dim : NULL,
                           j->n_rows, j->sdim, j->draws + (size_t)ci * j->n_rows * j->sdim,
                           j->stats + (size_t)ci * j->n_rows * NB200_NSTAT,
                           j->gradients ? j->gradients + (size_t)ci * j->n_rows * j->sdim : NULL,
                           j->mminv ? j->mminv + (size_t)ci * j->n_rows * j->sdim : NULL,
                           j->divs ? j->divs + (size_t)ci * j->n_rows * 4 * j->sdim : NULL,
                           j->eigvals ? j->eigvals + (size_t)ci * j->n_rows * j->st->mass_matrix_max_rank
                                      : NULL, &s);
        atomic_fetch_add(&j->steps, s);
        if (rc != 0) atomic_store(&j->err, rc);
    }
    return NULL;
}

/* one chain per task, n_threads host threads — the shape of nuts-rs's rayon
 * pool with `cores` workers (src/wrapper.rs:977, sample.py:1061-1070) */
int oracle_sample(const nb200_settings *st, nb200_logp_fn logp, const void *user_data,
                  uint64_t dim, uint64_t n_chains, uint64_t chain_id_offset, int n_threads,
                  const double *q0, const double *init_mean, const double *z_tape,
                  double *draws, double *stats, double *gradients, double *mass_matrix_inv,
                  uint64_t *total_steps) {
    return oracle_sample_ex(st, logp, user_data, dim, n_chains, chain_id_offset, n_threads, q0,
                            init_mean, z_tape, draws, stats, gradients, mass_matrix_inv, NULL,
                            total_steps);
}

int oracle_sample_ex(const nb200_settings *st, nb200_logp_fn logp, const void *user_data,
                     uint64_t dim, uint64_t n_chains, uint64_t chain_id_offset, int n_threads,
                     const double *q0, const double *init_mean, const double *z_tape,
                     double *draws, double *stats, double *gradients, double *mass_matrix_inv,
                     double *divergences, uint64_t *total_steps) {
    return oracle_sample_lr(st, logp, user_data, dim, n_chains, chain_id_offset, n_threads, q0,
                            init_mean, z_tape, draws, stats, gradients, mass_matrix_inv,
                            divergences, NULL, total_steps);
}

int oracle_sample_lr(const nb200_settings *st, nb200_logp_fn logp, const void *user_data,
                     uint64_t dim, uint64_t n_chains, uint64_t chain_id_offset, int n_threads,
                     const double *q0, const double *init_mean, const double *z_tape,
                     double *draws, double *stats, double *gradients, double *mass_matrix_inv,
                     double *divergences, double *eigvals, uint64_t *total_steps) {
    Job j;
    memset(&j, 0, sizeof(j));
    j.st = st; j.logp = logp; j.ud = user_data; j.dim = dim; j.n_chains = n_chains;
    j.chain_id_offset = chain_id_offset; j.q0 = q0; j.init_mean = init_mean; j.z_tape = z_tape;
    j.draws = draws; j.stats = stats; j.gradients = gradients; j.mminv = mass_matrix_inv;
    j.divs = divergences;
    j.eigvals = eigvals;
    j.n_total = st->num_tune + st->num_draws;
    j.n_rows = st->save_warmup ? j.n_total : st->num_draws;
    j.sdim = (st->store_dims && st->store_dims < dim) ? st->store_dims : dim;
    atomic_init(&j.next, 0); atomic_init(&j.steps, 0); atomic_init(&j.err, 0);
    if (n_threads <= 0) n_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if ((uint64_t)n_threads > n_chains) n_threads = (int)n_chains;
    if (n_threads <= 1) {
        worker(&j);
    } else {
        pthread_t *th = (pthread_t *)calloc(n_threads, sizeof(pthread_t));
        for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, worker, &j);
        for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    if (total_steps) *total_steps = atomic_load(&j.steps);
    return atomic_load(&j.err);
}
