/*
 * nutpie_b200.h — C-ABI of the B200-native NUTS sampler core.
 *
 * This header is the drop-in boundary for the path BASELINE.json's north_star
 * names: everything nutpie (pymc-devs/nutpie @ d3ffb850) delegates to the
 * external crate nuts-rs 0.18.3 — chain scheduling, the NUTS transition,
 * the leapfrog integrator with a diagonal mass matrix, dual-averaging
 * step-size adaptation and the Welford draw/gradient variance estimators —
 * re-expressed as hand-written sm_100a CUDA kernels behind plain C entry
 * points.  Each entry point cites the reference interface it replaces
 * (paths relative to /root/reference).
 *
 *   - plain pointers and sizes only, no C++/torch types;
 *   - every function returns 0 on success, a negative NB200_E* code on
 *     failure, and leaves a message retrievable with nb200_last_error();
 *   - the library owns its device memory (cudaMalloc) and its stream; host
 *     buffers handed in are copied before the call returns unless stated.
 *
 * The reference-side binding (the `extern "C"` block a maintainer would add
 * to src/wrapper.rs in place of `use nuts_rs::{Sampler, ...}`) is shown in
 * INTEGRATION.md.
 */
#ifndef NUTPIE_B200_H
#define NUTPIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB200_ABI_VERSION 4

/* ---- error codes ------------------------------------------------------- */
#define NB200_OK 0
#define NB200_EINVAL (-1)   /* bad argument / unsupported option            */
#define NB200_ECUDA (-2)    /* CUDA runtime error (see nb200_last_error)    */
#define NB200_ESTATE (-3)   /* call not valid in the sampler's state        */
#define NB200_ELOGP (-4)    /* fatal (non-recoverable) logp error, cf.      */
                            /* src/pymc.rs:166-181 (rc < 0 is fatal)        */
#define NB200_EINIT (-5)    /* no finite initial point found for a chain    */
#define NB200_ECOMPILE (-6) /* NVRTC missing or custom density failed to build */
#define NB200_ETIMEOUT 1    /* nb200_sampler_wait: still running            */

/* ---- settings ----------------------------------------------------------
 * Field-for-field the subset of nuts_rs::DiagNutsSettings that nutpie's
 * PyNutsSettings::apply_update can reach for adaptation="diag"/"draw_diag"
 * (src/wrapper.rs:210-451, 563-620).  Defaults: nb200_settings_default().
 */
typedef struct nb200_settings {
    uint64_t seed;             /* wrapper.rs:453-458 (random if None)          */
    uint64_t num_tune;         /* wrapper.rs:398-402; default 400              */
    uint64_t num_draws;        /* wrapper.rs:408-412; default 1000             */
    uint32_t maxdepth;         /* wrapper.rs:565-568; default 10               */
    uint32_t mindepth;         /* wrapper.rs:569-572; default 0                */
    int32_t check_turning;     /* wrapper.rs:573-576; default 1                */
    int32_t store_gradient;    /* wrapper.rs:393-397                           */
    int32_t store_mass_matrix; /* wrapper.rs:271-282                           */
    int32_t use_grad_based_estimate; /* wrapper.rs:283-294; 1 = "diag"         */
    double max_energy_error;   /* wrapper.rs:443-447; default 1000             */
    /* step size (nuts_rs step-size settings; wrapper.rs:240-270, 347-392) */
    double initial_step;       /* default 0.1                                  */
    double target_accept;      /* default 0.8                                  */
    double max_step_size;      /* dual_average.max_step_size; default +inf     */
    double da_k;               /* 0.75                                         */
    double da_t0;              /* 10                                           */
    double da_gamma;           /* 0.05                                         */
    int32_t step_size_method;  /* 0 dual_average | 1 adam | 2 fixed              */
                               /* (wrapper.rs:347-392)                         */
    int32_t _pad0;
    double fixed_step_size;    /* used when step_size_method == 2              */
    /* mass-matrix schedule (nuts_rs EuclideanAdaptOptions) */
    double early_window;       /* 0.3 of num_tune                              */
    double step_size_window;   /* 0.15 of num_tune: final step-size-only part  */
    uint64_t mass_matrix_switch_freq;        /* 80 (wrapper.rs:295-306)        */
    uint64_t early_mass_matrix_switch_freq;  /* 10 (wrapper.rs:225-239)        */
    uint64_t mass_matrix_update_freq;        /* 1                              */
    /* initial positions (nuts_rs::Model::init_position; src/pymc.rs:505-534,
     * src/stan.rs:798-808, src/pyfunc.rs:535-569) when no explicit q0 given */
    int32_t init_kind;         /* 0 uniform(-r,r)+mean | 1 normal(0,1)+mean    */
    int32_t num_try_init;      /* sample.py:860-863; default 10                */
    double init_radius;        /* 2.0 (pyfunc) / 1.0 (PyMC jitter)             */
    /* trace thinning (ours; config 4 cannot store 82 GB of raw draws) */
    uint64_t store_dims;       /* store only the first store_dims coordinates  */
                               /* of each draw; 0 = all                        */
    int32_t save_warmup;       /* sample.py:831; 1 = keep tuning draws         */
    int32_t expand_draws;      /* 1 = each stored draw is the model's EXPANDED  */
                               /* vector (constrained value variables followed  */
                               /* by deterministics), computed on the device:   */
                               /* CpuLogpFunc::expand_vector, src/pymc.rs:217-  */
                               /* 286; row width = nb200_model_expanded_dim().  */
                               /* Ignored when store_dims thins the draws.      */
    int32_t store_divergences; /* wrapper.rs:438-442: divergence_start / _end /  */
                               /* _momentum / _start_gradient rows (NaN unless   */
                               /* the draw diverged)                             */
    int32_t adaptation;        /* 0 diag (PyNutsSettings::Diag) | 1 low_rank     */
                               /* (PyNutsSettings::LowRank, wrapper.rs:307-346)  */
    double adam_learning_rate; /* wrapper.rs:376-392; default 0.05               */
    double step_size_jitter;   /* wrapper.rs:393-407; 0 = off                    */
    double mass_matrix_eigval_cutoff; /* wrapper.rs:307-326; default 2.0         */
    double mass_matrix_gamma;  /* wrapper.rs:327-346; default 1e-5               */
    uint64_t mass_matrix_max_rank; /* ours: at most this many eigenpairs are kept */
                               /* (largest |log eigenvalue| first; nuts-rs keeps   */
                               /* all beyond the cutoff); default 32, <= dim       */
} nb200_settings;

/* ---- model density plug-in ---------------------------------------------
 * The reference's own plug-in ABI (src/pymc.rs:23-29, produced by
 * python/nutpie/compile_pymc.py:970-1006) is a HOST function pointer.  The
 * device engine cannot call it per leapfrog, so a model is either one of the
 * built-in device densities below (hand-written CUDA, host twin in oracle/)
 * or NB200_MODEL_CUSTOM: the DEVICE analogue of that plug-in ABI (SURVEY
 * §8b2 / §8f-3), CUDA source compiled at run time with NVRTC into the same
 * persistent sampler kernel.  The source must define, at global scope,
 *
 *   __device__ int nb200_user_logp(const nb200_group &grp, int dim,
 *                                  const double *q, double *grad,
 *                                  double *logp_partial, const double *data);
 *
 * It is called by ALL grp.nthreads threads that own the chain (tid =
 * grp.tid): q [dim] and grad [dim] live in shared memory; every grad[i] must
 * be written by exactly one thread; *logp_partial receives this thread's
 * share of the log-density (the engine sums the shares in a fixed order);
 * grp.sum(x) is a bit-reproducible sum over the group that every thread
 * must call, grp.sync() a barrier over the group, grp.scratch [n_user_scratch]
 * the chain's private shared-memory scratch (contents do not persist between
 * calls); data [n_user_data] is the model's read-only device copy of
 * `user_data`.  Return code as in the
 * reference ABI (src/pymc.rs:178): 0 ok, > 0 recoverable (the trajectory
 * diverges there); non-finite logp / gradient are detected by the engine.
 * A serial density is `if (grp.tid == 0) { ... }`.
 */
typedef int (*nb200_logp_fn)(size_t dim, const double *x, double *grad_out,
                             double *logp_out, const void *user_data);
typedef int (*nb200_expand_fn)(size_t dim, size_t expanded_dim, const double *x,
                               double *out, const void *user_data);

#define NB200_MODEL_NORMAL 1 /* logp = -1/2 sum ((x-mu)/sigma)^2 ; configs 1,4 */
#define NB200_MODEL_FUNNEL 2 /* Neal's funnel: x0~N(0,1), x[1:]~N(0,exp(x0))   */
#define NB200_MODEL_RADON 3  /* README.md:53-88 / notebooks/pytensor_logp.md    */
                             /* :57-88 hierarchical radon, plain-Normal raws,   */
                             /* dim = 2*n_county + 5                            */
#define NB200_MODEL_CUSTOM 4 /* run-time compiled CUDA source (see above)       */
#define NB200_MODEL_HOST 5   /* the reference's own HOST plug-in, honoured bit   */
                             /* for bit: host_logp has the signature of          */
                             /* RawLogpFunc (src/pymc.rs:23-29, produced by      */
                             /* compile_pymc.py:970-1006), host_expand that of   */
                             /* RawExpandFunc (src/pymc.rs:31-37).  The sampler  */
                             /* kernel posts each position to a mailbox in       */
                             /* mapped pinned memory and a pool of host threads  */
                             /* calls the pointer (SURVEY.md §8f-3 fallback):    */
                             /* rc > 0 => the trajectory diverges there, rc < 0  */
                             /* => the sampler stops with NB200_ELOGP and keeps  */
                             /* its partial trace (src/pymc.rs:166-181).         */

typedef struct nb200_model_desc {
    int32_t kind;
    int32_t _pad;
    uint64_t dim;
    double mu, sigma;        /* NB200_MODEL_NORMAL                             */
    int32_t n_obs, n_county; /* NB200_MODEL_RADON                              */
    const double *y;         /* [n_obs] log_radon                              */
    const int32_t *county;   /* [n_obs] county index, any order                */
    const uint8_t *floor;    /* [n_obs] 0/1                                    */
    const char *cuda_source; /* NB200_MODEL_CUSTOM: NUL-terminated CUDA C++     */
    const double *user_data; /* NB200_MODEL_CUSTOM: [n_user_data] host doubles  */
    uint64_t n_user_data;
    uint64_t n_user_scratch; /* NB200_MODEL_CUSTOM: doubles of per-chain shared    */
                             /* memory handed to the density as grp.scratch     */
    nb200_logp_fn host_logp;     /* NB200_MODEL_HOST: LogpFunc.func               */
    const void *host_user_data;  /* LogpFunc.user_data_ptr (read-only, shared by  */
                                 /* all host threads: src/pymc.rs:47-48)          */
    nb200_expand_fn host_expand; /* ExpandFunc.func or NULL (draws stay           */
                                 /* unconstrained)                                */
    const void *host_expand_user_data;
    uint64_t host_expanded_dim;  /* ExpandFunc.expanded_dim                       */
    int32_t host_threads;        /* threads calling host_logp; 0 = all cores the  */
                                 /* process may use (the role of `cores`,         */
                                 /* src/wrapper.rs:977)                           */
    int32_t _pad2;
} nb200_model_desc;

/* ---- per-draw sampler statistics ---------------------------------------
 * One row of NB200_NSTAT doubles per (chain, draw); integers are stored
 * exactly as doubles.  Names follow nutpie's sample_stats
 * (docs/sample-stats.qmd:47-57, python/nutpie/sample.py:83).
 */
enum {
    NB200_STAT_DEPTH = 0,
    NB200_STAT_MAXDEPTH_REACHED = 1,
    NB200_STAT_INDEX_IN_TRAJECTORY = 2,
    NB200_STAT_LOGP = 3,
    NB200_STAT_ENERGY = 4,
    NB200_STAT_ENERGY_ERROR = 5,
    NB200_STAT_DIVERGING = 6,
    NB200_STAT_STEP_SIZE = 7,
    NB200_STAT_STEP_SIZE_BAR = 8,
    NB200_STAT_N_STEPS = 9,
    NB200_STAT_MEAN_TREE_ACCEPT = 10,
    NB200_STAT_MEAN_TREE_ACCEPT_SYM = 11,
    NB200_STAT_TUNING = 12,
    NB200_STAT_DRAW = 13,
    NB200_STAT_CHAIN = 14,
    NB200_STAT_RESERVED = 15,
    NB200_NSTAT = 16
};

/* src/wrapper.rs:38-104 (PyChainProgress) */
typedef struct nb200_progress {
    uint64_t finished_draws;
    uint64_t total_draws;
    uint64_t divergences;
    uint64_t latest_num_steps;
    uint64_t total_num_steps;
    double step_size;
    int32_t tuning;
    int32_t started;
} nb200_progress;

/* View into the library's pinned host copy of the trace.  Valid until the
 * next nb200_sampler_trace() call or nb200_sampler_destroy().
 * Replaces PySampler::{inspect,take_results,abort} → Vec<ArrowTrace>
 * (src/wrapper.rs:1332-1456).  Layout is ROW-major: element (row r, chain c)
 * starts at (r * n_chains + c) * width, so the rows that every chain has
 * finished are one contiguous block (they are streamed to the host with linear
 * copies while sampling runs); chain c's (posterior, sample_stats) RecordBatch
 * pair of src/wrapper.rs:1477-1494 is the strided slice [:, c, :]. */
typedef struct nb200_trace_view {
    uint64_t n_chains;
    uint64_t n_rows;         /* rows allocated per chain (tune+draws, or draws) */
    uint64_t dim;            /* model dimension                                 */
    uint64_t store_dims;     /* coordinates stored per draw                     */
    const double *draws;     /* [n_rows][n_chains][store_dims]                  */
    const double *stats;     /* [n_rows][n_chains][NB200_NSTAT]                 */
    const double *gradients; /* [n_rows][n_chains][dim or store_dims] or NULL   */
    const double *mass_matrix_inv; /* same shape as gradients, or NULL          */
    const uint64_t *rows_filled;   /* [n_chains] rows valid so far              */
} nb200_trace_view;

typedef struct nb200_sampler nb200_sampler;

/* ---- library ------------------------------------------------------------ */
int nb200_abi_version(void);
const char *nb200_last_error(void);
int nb200_device_count(void);
/* width of an expanded draw (ExpandFunc's expanded_dim, src/pymc.rs:78-94); 0 on error */
uint64_t nb200_model_expanded_dim(const nb200_model_desc *model);
/* nuts_rs::DiagNutsSettings::default() as surfaced by
 * PyNutsSettings::new_diag (src/wrapper.rs:525-533). */
void nb200_settings_default(nb200_settings *out);

/* ---- sampler life cycle -------------------------------------------------
 * nb200_sampler_create  ⇔ nuts_rs::Sampler::new(model, settings, storage,
 *                          cores, callback)        src/wrapper.rs:977-1085
 *   n_chains chains are created on `device`; chain c uses the random stream
 *   of GLOBAL chain id chain_id_offset + c, so a run sharded over several
 *   GPUs/processes reproduces the single-GPU run chain for chain.
 *   q0: optional host array [n_chains][dim] of initial positions
 *   (Model::init_position, src/pymc.rs:505-534); NULL → drawn on device per
 *   settings.init_kind around init_mean (NULL → zeros).
 * The call allocates and initialises device state and returns; sampling
 * starts with nb200_sampler_start (non-blocking, like Sampler::new). */
nb200_sampler *nb200_sampler_create(const nb200_settings *settings,
                                    const nb200_model_desc *model,
                                    uint64_t n_chains, uint64_t chain_id_offset,
                                    int device, const double *q0,
                                    const double *init_mean);
int nb200_sampler_start(nb200_sampler *s);
/* SamplerWaitResult / wait_timeout (src/wrapper.rs:1099-1185): returns
 * NB200_OK when finished, NB200_ETIMEOUT when still running after
 * timeout_seconds (<0 = forever), <0 on sampler error. */
int nb200_sampler_wait(nb200_sampler *s, double timeout_seconds);
/* progress callback payload (src/wrapper.rs:38-104); out has n_chains rows */
int nb200_sampler_progress(nb200_sampler *s, nb200_progress *out);
int nb200_sampler_is_finished(nb200_sampler *s); /* wrapper.rs:1252-1261 */
int nb200_sampler_pause(nb200_sampler *s);       /* wrapper.rs:1263-1282 */
int nb200_sampler_resume(nb200_sampler *s);      /* wrapper.rs:1284-1303 */
int nb200_sampler_abort(nb200_sampler *s);       /* wrapper.rs:1332-1365 */
/* inspect()/take_results(): copy the trace so far to pinned host memory
 * (src/wrapper.rs:1401-1456).  Allowed while running (snapshot). */
int nb200_sampler_trace(nb200_sampler *s, nb200_trace_view *out);
/* Copy the trace to caller-provided host buffers instead (may be NULL to
 * skip one); used by the end-to-end path to land draws directly in numpy /
 * Arrow buffers. */
int nb200_sampler_trace_into(nb200_sampler *s, double *draws, double *stats,
                             double *gradients, double *mass_matrix_inv,
                             uint64_t *rows_filled);
/* Register host buffers ([n_rows][n_chains][width] and [n_rows][n_chains][NB200_NSTAT],
 * ideally pinned) BEFORE start: nb200_sampler_wait then streams finished rows into them
 * while the kernel is still sampling, so the D2H copy of the trace overlaps the run;
 * nb200_sampler_trace_into with the same pointers afterwards copies nothing twice. */
int nb200_sampler_set_trace_target(nb200_sampler *s, double *draws, size_t draws_bytes,
                                   double *stats, size_t stats_bytes);
/* The same with ROW-STRIDED targets: row r of the draws lands at draws + r * draws_row_stride
 * (doubles; >= n_chains * width), likewise the stats.  The shards of one multi-GPU job — the
 * role of `cores` in src/wrapper.rs:977-1085 — write their chains into their own columns of the
 * job's single [row][chain][width] array, so the result needs no concatenation. */
int nb200_sampler_set_trace_target_strided(nb200_sampler *s, double *draws, size_t draws_row_stride,
                                           double *stats, size_t stats_row_stride);
/* exact sizes in bytes the two buffers above must have */
int nb200_sampler_trace_bytes(nb200_sampler *s, size_t *draws_bytes, size_t *stats_bytes);
int nb200_sampler_destroy(nb200_sampler *s);

/* ---- measurement hooks --------------------------------------------------
 * Device time (CUDA events recorded on the sampler's own stream around the
 * sampling kernel launches) and launch count, for bench.py. */
double nb200_sampler_kernel_ms(nb200_sampler *s);
uint64_t nb200_sampler_launch_count(nb200_sampler *s);
/* raw device pointers (for torch.distributed gathers without a host hop) */
int nb200_sampler_device_buffers(nb200_sampler *s, void **draws, void **stats);
/* launch geometry chosen for this sampler (threads per chain, block, grid) */
int nb200_sampler_geometry(nb200_sampler *s, int32_t *threads_per_chain,
                           int32_t *block, int32_t *grid);
/* override threads per chain (0 = auto; 32..1024, power of two) and chains
 * per CTA (0 = auto; only used when threads per chain == 32) for samplers
 * created afterwards; testing/tuning */
void nb200_set_threads_per_chain(int32_t t);
void nb200_set_chains_per_block(int32_t c);
/* force the number of pool slots kept in shared memory (-1 = auto) */
void nb200_set_smem_slots(int32_t n);
int nb200_sampler_smem(nb200_sampler *s, int32_t *smem_slots, int32_t *bytes_per_chain);
/* 0 = always use the run-time-trip-count kernels (testing); 1 = auto */
void nb200_set_unroll(int32_t on);
/* streaming leapfrog (256 threads per chain), bit mask: 1 = inputs read through bulk-copy
 * staging, 2 = successive passes sweep the dimensions in alternating directions (the tail a
 * pass wrote is the first thing the next one reads), 4 = L2 eviction hints on the bulk copies */
void nb200_set_stage_loads(int32_t mode);
/* 1 (default) = densities that gather across dimensions run two warps per chain: an integrator
 * warp producing leaves back to back and a tree warp consuming them (bit-identical traces);
 * 0 = one warp does both.  nb200_sampler_is_pipelined reports what a sampler got. */
void nb200_set_pipeline(int32_t on);
int nb200_sampler_is_pipelined(nb200_sampler *s);
/* limit the draws one kernel launch may advance each chain by (0 = run to the
 * end in one persistent launch); the host relaunches until done */
int nb200_sampler_set_draws_per_launch(nb200_sampler *s, uint64_t n);
/* tests: replace the momentum stream by a host tape [n_chains][tune+draws][dim]
 * of standard normals; must precede start() */
int nb200_sampler_set_z_tape(nb200_sampler *s, const double *z_tape);
/* pinned host memory for trace buffers handed to nb200_sampler_trace_into */
/* ---- run-time compiled densities ------------------------------------------
 * Compile `model->cuda_source` (kind NB200_MODEL_CUSTOM) for sm_100a into the
 * sampler kernel specialised for `threads_per_chain` (32..1024, power of two)
 * and `dims_per_thread` unrolled dimensions per thread (0 = run-time loop).
 * Needs libnvrtc but no GPU; nb200_sampler_create does the same lazily for
 * the geometry it picks.  On failure returns NB200_ECOMPILE and copies the
 * NVRTC log into log[0..log_len).  ⇔ the role of numba's cfunc compile in
 * python/nutpie/compile_pymc.py:970-1006.                                   */
int nb200_custom_model_compile(const nb200_model_desc *model,
                               int threads_per_chain, int dims_per_thread,
                               char *log, size_t log_len);

/* store_divergences: copy the divergence rows [n_rows][n_chains][4][width] (start location,
 * end location, start momentum, start gradient — nuts-rs DivergenceInfo as surfaced by
 * python/nutpie/sample.py:641-646; NaN for draws that did not diverge; width as `gradients`) */
int nb200_sampler_divergence_trace_into(nb200_sampler *s, double *divergences);
/* adaptation = low_rank with store_mass_matrix (the reference's mass_matrix_eigvals column,
 * tests/test_pymc.py:116-131; mass_matrix_stds arrives in the mass_matrix_inv rows): copy the
 * eigenvalue rows [n_rows][n_chains][max_rank], NaN beyond the rank in use.  *max_rank
 * receives the row width; eigvals may be NULL to query it. */
int nb200_sampler_eigvals_trace_into(nb200_sampler *s, double *eigvals, uint64_t *max_rank);
/* CpuLogpFunc::expand_vector (src/pymc.rs:217-286) over a finished trace of a
 * NB200_MODEL_HOST model: out[i] = expand(q[i]) for n rows, on n_threads host threads
 * (0 = all).  q rows are q_stride doubles apart, out rows expanded_dim.  Returns the first
 * non-zero return code of `fn` as NB200_ELOGP. */
int nb200_host_expand_rows(nb200_expand_fn fn, const void *user_data, size_t dim,
                           size_t expanded_dim, uint64_t n, const double *q,
                           size_t q_stride, double *out, int n_threads);

void *nb200_host_alloc(size_t bytes);
void nb200_host_free(void *p);

/* ---- component entry points (parity tests at the nuts-rs Math seam) -----
 * nb200_logp_grad  ⇔ CpuLogpFunc::logp (src/pymc.rs:197-215): evaluate the
 *   device density for n points; rc[i] follows the reference convention
 *   (0 ok, 3 non-finite grad, 4 non-finite logp; compile_pymc.py:996-999).
 * nb200_leapfrog   ⇔ one nuts-rs leapfrog with a diagonal mass matrix for n
 *   independent states; all arrays host [n][dim]
 *   except eps [n], dir [n], and outputs logp/kinetic [n].
 */
int nb200_logp_grad(const nb200_model_desc *model, int device, uint64_t n,
                    const double *q, double *logp, double *grad, int32_t *rc);
int nb200_leapfrog(const nb200_model_desc *model, int device, uint64_t n,
                   const double *q, const double *p, const double *g,
                   const double *var, const double *p_sum, const double *eps,
                   const int32_t *dir, const int64_t *idx, double *q_out,
                   double *p_out, double *g_out, double *p_sum_out,
                   double *logp_out, double *kinetic_out, int32_t *rc);
/* low-rank metric at the component seam (oracle twin: oracle_lowrank_update / _velocity /
 * _momentum): refresh the metric from a window of n draws and gradients [n][dim] on the device,
 * then v_out[v] = M^-1 p[v] and momentum_out[v] = M^1/2 z[v] for n_vec vectors [n_vec][dim].
 * stds_out [dim], vals_out [max_rank], vecs_out [max_rank][dim], *rank_out may be NULL. */
int nb200_lowrank_component(int device, uint64_t dim, uint64_t n, const double *draws,
                            const double *grads, double gamma, double cutoff,
                            uint64_t max_rank, uint64_t n_vec, const double *p, double *v_out,
                            const double *z, double *momentum_out, double *stds_out,
                            double *vals_out, double *vecs_out, uint64_t *rank_out);

#ifdef __cplusplus
}
#endif
#endif /* NUTPIE_B200_H */
