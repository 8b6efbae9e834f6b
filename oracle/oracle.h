/*
 * oracle/oracle.h — TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement of the sampler core nutpie delegates to nuts-rs 0.18.3
 * (Cargo.toml:24, Cargo.lock:2295-2298).  nuts-rs is a crates.io dependency
 * that is NOT vendored under /root/reference and there is no Rust toolchain
 * here, so this file restates its published algorithm (SURVEY.md Appendix A;
 * Hoffman & Gelman 2014; Betancourt 2017 multinomial NUTS; Seyboldt et al.
 * arXiv:2603.18845 for the gradient-based diagonal mass matrix) and anchors
 * it on nutpie's own call sites and in-tree pins.
 *
 *   PARITY STATUS: bit-level parity with the reference is UNPINNED (the
 *   reference's seeded streams in tests/reference/ (*.txt) need rand 0.10
 *   ChaCha8 + rand_distr ziggurat + PyMC's seeded init, none reachable
 *   here).  Pinned instead: component known answers
 *   (python/nutpie/normalizing_flow.py:1905-1915 for the mass-matrix
 *   estimate; the dual-averaging recursion; closed-form leapfrog / U-turn),
 *   the determinism contract of tests/test_stan.py:67-101,282-302, the
 *   analytic-posterior checks of tests/test_pymc.py:397-416, and the
 *   distribution of tests/reference/test_deterministic_sampling_numba.txt.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>
#include <stdint.h>

#include "../include/nutpie_b200.h" /* nb200_settings, NB200_STAT_* (interface only) */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_normal_data {
    double mu, sigma;
} oracle_normal_data;

typedef struct oracle_radon_data {
    int32_t n_obs, n_county;
    const double *y;
    const int32_t *county;
    const uint8_t *floor;
} oracle_radon_data;

/* densities in the reference plug-in ABI (src/pymc.rs:23-29) */
int oracle_logp_normal(size_t dim, const double *x, double *grad, double *logp, const void *ud);
int oracle_logp_halfnormal(size_t dim, const double *x, double *grad, double *logp, const void *ud);
int oracle_logp_funnel(size_t dim, const double *x, double *grad, double *logp, const void *ud);
int oracle_logp_radon(size_t dim, const double *x, double *grad, double *logp, const void *ud);
int oracle_logp_logreg(size_t dim, const double *x, double *grad, double *logp, const void *ud);
int oracle_expand_radon(size_t dim, size_t expanded_dim, const double *x, double *out,
                        const void *ud);

/* Run n_chains chains (global ids chain_id_offset..), one chain per task on
 * n_threads host threads (<=0: all).  Outputs are host arrays laid out exactly
 * like nb200_trace_view: draws [n_chains][n_rows][store_dims], stats
 * [n_chains][n_rows][NB200_NSTAT], optional gradients / mass_matrix_inv.
 * q0 / init_mean as in nb200_sampler_create.  z_tape: optional
 * [n_chains][num_tune+num_draws][dim] standard normals replacing the
 * RNG_MOMENTUM stream (tape-driven parity).
 * Returns 0, or a negative NB200_E* code. */
int oracle_sample(const nb200_settings *settings, nb200_logp_fn logp, const void *user_data,
                  uint64_t dim, uint64_t n_chains, uint64_t chain_id_offset, int n_threads,
                  const double *q0, const double *init_mean, const double *z_tape,
                  double *draws, double *stats, double *gradients, double *mass_matrix_inv,
                  uint64_t *total_steps);

/* the same with the divergence rows of store_divergences:
 * divergences [n_chains][n_rows][4][store_dims] (NaN unless the draw diverged) */
int oracle_sample_ex(const nb200_settings *settings, nb200_logp_fn logp, const void *user_data,
                     uint64_t dim, uint64_t n_chains, uint64_t chain_id_offset, int n_threads,
                     const double *q0, const double *init_mean, const double *z_tape,
                     double *draws, double *stats, double *gradients, double *mass_matrix_inv,
                     double *divergences, uint64_t *total_steps);

/* the same with adaptation = low_rank traces: with store_mass_matrix the mass_matrix_inv rows
 * hold mass_matrix_stds, and eigvals [n_chains][n_rows][mass_matrix_max_rank] the kept
 * eigenvalues (NaN-padded) — the reference's mass_matrix_stds / mass_matrix_eigvals columns */
int oracle_sample_lr(const nb200_settings *settings, nb200_logp_fn logp, const void *user_data,
                     uint64_t dim, uint64_t n_chains, uint64_t chain_id_offset, int n_threads,
                     const double *q0, const double *init_mean, const double *z_tape,
                     double *draws, double *stats, double *gradients, double *mass_matrix_inv,
                     double *divergences, double *eigvals, uint64_t *total_steps);

/* low-rank metric components (oracle/lowrank.c): estimate from a window of n draws / gradients
 * [n][dim]; v = M^-1 p; p = M^1/2 z */
int oracle_lowrank_update(size_t dim, size_t n, const double *draws, const double *grads,
                          double gamma, double cutoff, size_t max_rank, double *stds,
                          double *vals, double *vecs, size_t *rank_out);
void oracle_lowrank_velocity(size_t dim, const double *stds, size_t k, const double *vals,
                             const double *vecs, const double *p, double *v);
void oracle_lowrank_momentum(size_t dim, const double *stds, size_t k, const double *vals,
                             const double *vecs, const double *z, double *p);

/* --- component entry points (known-answer tests at the nuts-rs Math seam) --- */
/* one leapfrog, SURVEY Appendix A.2 */
int oracle_leapfrog(nb200_logp_fn logp, const void *ud, size_t dim, const double *q,
                    const double *p, const double *g, const double *var, const double *p_sum,
                    double eps, int dir, int64_t idx, double *q_out, double *p_out,
                    double *g_out, double *p_sum_out, double *logp_out, double *kinetic_out);
/* U-turn criterion between two states of one trajectory, Appendix A.3 */
int oracle_is_turning(size_t dim, int64_t idx1, const double *p1, const double *psum1,
                      int64_t idx2, const double *p2, const double *psum2, const double *var);
/* dual averaging: state = {log_step, log_step_adapted, hbar, mu, count} */
void oracle_dual_average_init(double state[5], double initial_step);
void oracle_dual_average_advance(double state[5], double accept_stat, double target, double k,
                                 double t0, double gamma);
/* Adam step-size rule on the same 5-slot state (hbar = m, mu = v, count = t) */
void oracle_adam_advance(double state[5], double accept_stat, double target, double lr);
/* running variance (Welford): mean[D], m2[D], *count */
void oracle_welford_add(size_t dim, double *mean, double *m2, uint64_t *count, const double *x);
/* mass-matrix refresh from the two estimators (grad-based) or draws only */
void oracle_mass_matrix_update(size_t dim, int use_grad, const double *m2_draw,
                               const double *m2_grad, uint64_t count, double *var_out);
void oracle_mass_matrix_init(size_t dim, const double *grad, double *var_out);
/* RNG stream, for cross-checks against the device implementation */
void oracle_rng_words(uint64_t seed, uint32_t chain, uint32_t draw, uint32_t purpose,
                      uint32_t index, uint64_t out[2]);
void oracle_rng_normals(uint64_t seed, uint32_t chain, uint32_t draw, uint32_t purpose,
                        size_t dim, double *z);

#ifdef __cplusplus
}
#endif
#endif
