// nuts_core.cuh — the per-chain NUTS engine (transition + adaptation), written
// once against a thread-group policy `G`, a density `M` and a compile-time trip
// count `NIT` of the per-dimension loops.
//
// What it replaces (nuts-rs 0.18.3 behind nuts_rs::Sampler::new,
// src/wrapper.rs:977-1085; semantics per SURVEY.md Appendix A):
//   leapfrog()            — Euclidean Hamiltonian with diagonal mass matrix (A.2)
//   is_turning()          — U-turn criterion on momentum prefix sums (A.3)
//   transition()          — tree doubling, multinomial + biased-progressive
//                           selection, divergence rule (A.4); ITERATIVE: the
//                           crate's recursion is unrolled onto a binary-counter
//                           stack of sub-trees, one entry per level
//   adapt()               — dual averaging, Welford draw+gradient estimators
//                           with foreground/background windows, mass-matrix
//                           refresh, initial step-size search (A.5)
//
// Memory plan (B200): every trajectory state is a SLOT of four vectors
// (q, p, grad, p_sum) in a per-chain pool in HBM; a leapfrog reads one slot and
// writes a fresh one exactly once (72 B per dimension per gradient evaluation,
// the algorithmic figure of SURVEY.md §8d) and all tree bookkeeping is done on
// slot indices, so merging sub-trees never copies a vector.  Per-slot scalars
// (index in trajectory, potential, kinetic energy) and the level stack live in
// shared memory.
//
// Code-size discipline: with one warp per chain the kernel is bound by instruction
// fetch (1-2 warps per scheduler walking ~3 K instructions per leapfrog), so the hot
// loop keeps ONE inlined copy of leapfrog() and of is_turning() (a rolled loop walks
// the pairs of a merge), one loop body per pass, rolled loops where unrolling buys
// nothing; the only other copies of leapfrog() are the two cold ones inside the
// step-size search.  The first version was 335 KB of SASS; every cut since measured
// as a speed-up (profiles/r1_sweep_radon_codesize.txt).
#pragma once
#include "../../include/nutpie_b200.h"
#include "group.cuh"
#include "lowrank.cuh"
#include "models.cuh"
#include "philox.cuh"
#include "portable.cuh"

// partners (besides the predecessor) whose U-turn pair is evaluated inside the streaming
// leapfrog pass; tunable at build time for register-pressure experiments
#ifndef NB200_MAX_FUSED
#define NB200_MAX_FUSED 1  // measured best for config 4: profiles/r1_sweep_fused_partners.txt
#endif

namespace nb200 {

// Streaming leapfrog staging (CTA-per-chain geometry with T = 128 or 256 threads): kStages
// buffers of one chunk each; a chunk is T double2 (one per thread) of every source vector,
// 4 source vectors + 2 per fused partner; one mbarrier per stage behind the buffers.
#ifndef NB200_STAGES
#define NB200_STAGES 4  // at 128 threads per chain; measured: profiles/r1_sweep_config4_variants.txt
#endif
constexpr int kStageVecs = 4 + 2 * NB200_MAX_FUSED;
// stages are sized so that four chains fit one SM's shared memory (<= 48 KB of staging each)
template <int T>
NB_HD constexpr int stage_count() {
    return T == 256 ? 2 : NB200_STAGES;
}
template <int T>
NB_HD constexpr int stage_buf_bytes() {
    return kStageVecs * T * 16;  // one stage
}
template <class M, int T>
NB_HD constexpr int stage_smem_bytes() {
    return (M::kElementwise && (T == 128 || T == 256))
               ? stage_count<T>() * stage_buf_bytes<T>() + ((8 * stage_count<T>() + 15) & ~15)
               : 0;
}

constexpr int kMaxSlots = 64;
constexpr int kMaxLevels = 20;
constexpr double kVarLower = 1e-20, kVarUpper = 1e20;

enum { VQ = 0, VP = 1, VG = 2, VS = 3, VV = 4 };  // q, p, grad, p_sum inside a slot; low-rank
                                                   // adaptation adds the velocity M^-1 p

// persistent per-chain scalar state (global memory; also the progress record)
struct ChainScalars {
    double step_size;
    double da_log_step, da_log_step_adapted, da_hbar, da_mu;
    unsigned long long da_count;
    unsigned long long cnt[2];   // sample counts of the two Welford sets
    unsigned long long last_update;
    unsigned long long draw;     // next draw index == finished draws
    unsigned long long total_steps;
    unsigned long long divergences;
    unsigned long long latest_n_steps;
    unsigned long long published;  // rows < published are complete in HBM (fenced): safe to copy out
    double cur_U;                // potential energy (-logp) of the current point
    int fg_sel;                  // which Welford set is the foreground
    int has_initial_mm;
    int cur_slot;
    int status;                  // 0 not started, 1 running, 2 finished, <0 NB200_E*
    int sweep_rev;               // direction of the last streaming pass (a relaunch must continue
                                 // the alternation: the sweep order fixes the order of the sums)
    int lr_k, lr_len, lr_split, lr_head;  // low-rank adaptation: rank in use, window deque
};

// Producer / consumer pipeline of one chain (two warps, see ChainCtx::producer_main): a ring of
// kPipeDepth leaves in flight.  The CONSUMER (tree warp) grants entry e a destination slot and
// arrives on empty[e]; the PRODUCER (integrator warp) runs the leapfrog into that slot, writes
// the leaf's scalars and arrives on full[e].  dst = -1 is the end-of-draw marker.
constexpr int kPipeDepth = 4;
struct PipeShared {
    unsigned long long full[kPipeDepth], empty[kPipeDepth];  // mbarriers (count 1)
    unsigned long long cmd;                                   // mbarrier: next command posted
    double de[kPipeDepth];   // energy error of the leaf
    double step_size, E0;    // command: step size of the draw; reply: energy of its first point
    int dst[kPipeDepth], rc[kPipeDepth], l0[kPipeDepth];
    int cur;                 // command: slot of the current point
    unsigned draw;           // command: draw index
    int quit;                // command: the chain is done for this launch
    int stop;                // consumer -> producer: skip the leaves still granted
};

// per-chain shared-memory scalars
struct ChainShared {
    double U[kMaxSlots];
    double K[kMaxSlots];
    double lvLS[kMaxLevels];
    int idx[kMaxSlots];
    int lvL[kMaxLevels], lvR[kMaxLevels], lvD[kMaxLevels];
    int stop;  // the host's stop flag as read by thread 0 (one decision for the whole group)
    PipeShared pipe;
};

template <class M>
struct KParams {
    nb200_settings st;
    typename M::Data mdata;
    int D, Dp, NS;
    int smem_slots;    // pool slots 0..smem_slots-1 live in shared memory (hot tier)
    int var_in_smem;   // working copy of the mass matrix in shared memory
    int expand;        // stored draws are expanded vectors (sdim = expanded dimension)
    unsigned long long n_chains, chain_id_offset;
    unsigned long long n_rows, sdim, n_total;
    unsigned long long gdim;  // row width of the gradient / mass-matrix traces
    unsigned long long max_draws_per_launch;  // 0 = run to the end
    double* pool;      // [n_chains][NS][4][Dp]
    double* var;       // [n_chains][Dp]        diagonal of M^-1
    double* welford;   // [n_chains][2][4][Dp]  (mean_q, m2_q, mean_g, m2_g) x 2 sets
    ChainScalars* sc;  // [n_chains]
    double* draws;     // [n_rows][n_chains][sdim]
    double* stats;     // [n_rows][n_chains][NB200_NSTAT]
    double* grads;     // optional, like draws
    double* mminv;     // optional, like draws
    double* divs;      // optional [n_rows][n_chains][4][gdim]: divergence start / end location,
                       // start momentum, start gradient (store_divergences); NaN-filled by the host
    const double* q0;        // optional [n_chains][D]
    const double* init_mean; // optional [D]
    const double* z_tape;    // optional [n_chains][n_total][D] (tests)
    const volatile int* stop_flag;
    // low-rank adaptation (st.adaptation == 1; lowrank.cuh): per-chain metric, window and scratch
    double* lr_stds;   // [n_chains][Dp]
    double* lr_vals;   // [n_chains][lr_max_rank]
    double* lr_vecs;   // [n_chains][lr_max_rank][Dp]
    double* lr_coef;   // [n_chains][lr_max_rank]
    double* lr_win;    // [n_chains][lr_cap][2][Dp]
    double* lr_mat;    // [n_chains][2][D][Dp]
    double* lr_cols;   // [n_chains][6][Dp]
    double* eigvals;   // optional trace [n_rows][n_chains][lr_max_rank] (store_mass_matrix), NaN-padded
    int lr_cap, lr_max_rank;
    // streaming leapfrog (bit mask): 1 = inputs come through bulk-copy staging; 2 = successive
    // passes sweep the dimensions in alternating directions, so the tail a pass has just written
    // is the first thing the next pass reads (still in L2); 4 = the bulk copies carry L2
    // eviction hints (consumed state leaves first, the mass matrix stays)
    int stage_loads;
};

struct SampleInfo {
    int depth, diverging, maxdepth_reached;
};

// log(exp(a) + exp(b)); one inlined exp / log1p pair for both orderings (the two-branch form of
// the oracle computes the same values with twice the code in the merge loop)
NB_HD double nb_logaddexp(double a, double b) {
    if (a == b) return a + 0.69314718055994530941723212145818;
    const double diff = a - b;
    if (diff != diff) return diff;  // NaN
    const double hi = diff > 0 ? a : b;
    return hi + log1p(exp(-fabs(diff)));
}

#ifdef __CUDACC__
// Acceptance statistics of up to 32 parked leaves (ChainCtx::flush_acc): lane k holds leaf k's
// energy error; returns (sum of min(1, w), sum of 2 min(1, w) / (1 + w)) added to the running
// sums in leaf order.  One out-of-line copy: it runs once per transition, and inlining its exp /
// division at every call site only grows the hot loop's instruction footprint.
static __device__ __noinline__ double2 nb_flush_parked(double de, unsigned n, double sum, double sym) {
    double a = 0.0, sy = 0.0;
    if ((threadIdx.x & 31u) < n) {
        const double w = exp(-de);
        a = w < 1.0 ? w : 1.0;
        sy = 2.0 * a / (1.0 + w);
    }
    __syncwarp();
    for (unsigned k = 0; k < n; ++k) {
        sum += __shfl_sync(0xffffffffu, a, (int)k);
        sym += __shfl_sync(0xffffffffu, sy, (int)k);
    }
    return make_double2(sum, sym);
}
#endif

#ifdef __CUDACC__
// The uniforms of a transition's merges (Philox keyed by the merge's sequence number) do not
// depend on anything the transition computes: lane k draws the one for merge base + k, so a warp
// runs the generator once per 32 merges instead of once per merge (ChainCtx::merge_uniform).
// Out of line: it runs rarely and its 80 instructions stay out of the hot loop.
static __device__ __noinline__ double nb_refill_merge_uniforms(uint64_t seed, uint32_t chain, uint32_t draw,
                                                               uint32_t base) {
    uint64_t a, b;
    rng_u64x2(seed, chain, draw, RNG_MERGE, base + (threadIdx.x & 31u), a, b);
    return rng_u01(a);
}
#endif

// NIT > 0: every per-dimension loop runs exactly NIT predicated iterations
// (NIT * group size >= D), fully unrolled so independent loads overlap;
// NIT == 0: run-time trip count (any D).
template <class M, class G, int NIT = 0, bool LR = false>
struct ChainCtx {
    static_assert(!LR || NIT == 0, "the low-rank engine runs the run-time loops");
    static constexpr int kVecs = LR ? 5 : 4;  // vectors per pool slot
    LrState lr;  // low-rank adaptation only
    G g;
    const KParams<M>* P;
    typename M::Data md;  // density data (a per-CTA copy may point into shared memory)
    ChainShared* sh;
    double* msm;  // model scratch (shared)
    // "front": shared-memory copy (q, p, grad, p_sum) of the newest leaf of the running
    // trajectory.  Densities that gather across dimensions integrate IN PLACE on it
    // with plain LDS/STS and only write each new state out to its pool slot; the
    // U-turn checks read the newest leaf from it.  front_slot = pool slot it mirrors.
    double* front;
    int front_slot;
    int D, Dp, NS;
    unsigned long long chain_local;
    uint32_t chain_gid;
    double *pool, *var, *wf;
    unsigned char* stage;   // bulk-copy staging buffers + mbarriers (streaming geometry only)
    unsigned stage_phase;   // parity bit per stage barrier
    bool sweep_rev;         // direction of the last streaming leapfrog over the dimensions
    double* spool;     // shared-memory tier of the pool (slots < smem_slots)
    double* varg;      // persistent (global) copy of the mass matrix; var may alias it
    int smem_slots;
    // hamiltonian
    double step_size, E0;
    // collectors of the running transition (AcceptanceRateCollector)
    double acc_sum, acc_sym;
    uint32_t acc_count, n_merge, draw;
    double last_de;
    // Warp-per-chain geometry: the acceptance statistics of a transition's leaves are evaluated
    // lane-parallel.  Leaf n parks its energy error in lane n % 32; every 32 leaves and at the
    // end of the transition each lane evaluates exp / the symmetric ratio for ITS leaf and the
    // values are added in leaf order (same sums, bit for bit, as adding them leaf by leaf — but
    // one exp + one division per 32 leaves instead of per leaf on the warp's critical path).
    double de_parked;
    unsigned n_parked;
    bool defer_acc;
    double u_parked;   // merge uniforms drawn ahead, one per lane (warp-per-chain geometry)
    uint32_t u_base;   // u_parked of lane k belongs to merge u_base + k
    bool l0_turn;          // U-turn verdict between src and the new leaf, fused into the leapfrog
    unsigned fused_bits;   // verdicts of the other pairs planned for this leapfrog (bit c = plist[c])
    // tree bookkeeping (slot ids; -1 = none)
    int mL, mR, mD, tL, tR, tD;
    unsigned long long lv_live;  // slots referenced by the parked sub-trees of the level stack
    // strategy state
    double da_log_step, da_log_step_adapted, da_hbar, da_mu;
    unsigned long long da_count;
    unsigned long long cnt0, cnt1, last_update, total_steps, divergences;
    int fg_sel, has_initial_mm;
    double last_mean, last_sym;
    uint32_t last_n_steps;
    int div_src, div_dst;  // slots of the leapfrog that diverged in the running transition
    // two-warp pipeline (device only): this warp's running leaf / grant / command counters,
    // the slots granted to the producer but not yet received back, collector switch
    bool piped = false, skip_acc = false;
#ifdef NB200_PIPE_PROFILE
    long long prof[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#ifdef __CUDA_ARCH__
#define NB_PROF_CLK() clock64()
#else
#define NB_PROF_CLK() 0ll
#endif
#define NB_PROF_T0() const long long prof_t0_ = NB_PROF_CLK()
#define NB_PROF_ADD(i) prof[i] += NB_PROF_CLK() - prof_t0_
#define NB_PROF_INC(i) prof[i] += 1
// consecutive regions of one function: MARK(i) charges the time since the previous mark to slot i
#define NB_PROF_MARK_INIT() long long prof_mark_ = piped ? 0 : NB_PROF_CLK()
#define NB_PROF_MARK(i)                                  \
    do {                                                 \
        if (!piped) {                                    \
            const long long now_ = NB_PROF_CLK();        \
            prof[i] += now_ - prof_mark_;                \
            prof_mark_ = now_;                           \
        }                                                \
    } while (0)
#else
#define NB_PROF_MARK_INIT()
#define NB_PROF_MARK(i)
#define NB_PROF_T0()
#define NB_PROF_ADD(i)
#define NB_PROF_INC(i)
#endif
    unsigned kcons = 0, kgrant = 0, ncmd = 0;
    unsigned long long reserved = 0;

    // f(i) for every dimension index owned by this thread
    template <class F>
    NB_HD void for_dims(F&& f) const {
        if constexpr (NIT > 0) {
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int i = g.tid + it * G::kThreads;
                // NIT = ceil(D / group size): only the last iteration can run past D
                if (it + 1 < NIT || i < D) f(i);
            }
        } else {
            for (int i = g.tid; i < D; i += g.size()) f(i);
        }
    }

    // Hot tier in shared memory: only the MOMENTUM half (p, p_sum) of the lowest
    // 2 * smem_slots pool slots.  That half is what the U-turn checks read from older
    // states; positions and gradients of older states are read back only when a
    // trajectory restarts from a tree end or a draw is written out, so they always
    // live in the HBM/L2 tier.  alloc() hands out the lowest free slot, so recent
    // leaves and low tree levels land in the hot tier.
    // (a chain's pool is at most 64 slots x 4 vectors: offsets fit 32 bits for any D < 2^23)
    NB_HD double* vec(int slot, int comp) const {
        if ((comp & 1) && slot < 2 * smem_slots)  // VP = 1, VS = 3
            return spool + (unsigned)((slot * 2 + (comp >> 1)) * Dp);
        return pool + (unsigned)((slot * kVecs + comp) * Dp);
    }
    NB_HD double* gvec(int slot, int comp) const { return pool + (unsigned)((slot * kVecs + comp) * Dp); }
    NB_HD const nb200_settings& st() const { return P->st; }

    // ------------------------------------------------------------ slot pool
    NB_HD int alloc() const {
        unsigned long long live = 0;
        if (mL >= 0) live |= 1ull << mL;
        if (mR >= 0) live |= 1ull << mR;
        if (mD >= 0) live |= 1ull << mD;
        if (tL >= 0) live |= 1ull << tL;
        if (tR >= 0) live |= 1ull << tR;
        if (tD >= 0) live |= 1ull << tD;
        live |= lv_live | reserved;
        const unsigned long long all = NS >= 64 ? ~0ull : ((1ull << NS) - 1ull);
        return nb_ffsll(~live & all) - 1;  // lowest free slot: keeps the hot set small
    }

    // ---------------------------------------------------------- density at a slot
    NB_HD double eval_logp(int slot, bool& bad) {
        const double* q = vec(slot, VQ);
        double* gr = vec(slot, VG);
        double lp;
        double flag[1] = {0.0};
        if constexpr (M::kElementwise) {
            double acc[2] = {0.0, 0.0};
            for_dims([&](int i) {
                double gn;
                acc[0] += M::term(md, i, q[i], gn);
                gr[i] = gn;
                if (!nb_isfinite(gn)) acc[1] += 1.0;
            });
            g.reduce(acc);
            lp = M::finish(md, acc[0], D);
            flag[0] = acc[1];
        } else {
            double* fq = front;
            double* fg = front + 2 * (size_t)Dp;
            for_dims([&](int i) { fq[i] = q[i]; });
            g.sync();
            lp = M::logp_grad(g, md, D, fq, fg, msm);
            g.sync();
            for_dims([&](int i) {
                const double gn = fg[i];
                gr[i] = gn;
                if (!nb_isfinite(gn)) flag[0] += 1.0;
            });
            g.reduce(flag);
            front_slot = -1;  // momentum part of the front is stale
        }
        bad = flag[0] > 0.0 || !nb_isfinite(lp);
        return lp;
    }

    // ------------------------------------------------- fused U-turn bookkeeping
    // The checks that pair the NEW leaf with an older state are evaluated inside the
    // leapfrog's own pass over the dimensions (the new p, p_sum are in registers there):
    //   * the pair (src, new) costs nothing extra — src's vectors are loaded anyway;
    //   * up to kMaxFused further partners cost two vector reads each instead of a
    //     separate five-vector pass plus its own reduction.
    // All verdicts share one reduction with logp / kinetic energy.
    static constexpr bool kFuseL0 = (M::kElementwise || NIT > 0) && !LR;  // not the run-time-loop front path
    // Measured (profiles/r1_sweep_*): planning + extra partner loads pay off only in the
    // streaming regime (large D, run-time loops), where each separate check is a pass over
    // HBM; for small D they cost more than the on-chip is_turning() they replace.
    static constexpr int kMaxFused = (M::kElementwise && NIT == 0 && !LR) ? NB200_MAX_FUSED : 0;
    static constexpr int kFusedDim = kMaxFused > 0 ? kMaxFused : 1;

    // The pair (x, new leaf N) in is_turning()'s terms: with (s, e) = the earlier / later of
    // the two, rho is one of three forms depending on where the pair sits relative to the
    // trajectory's origin — (S_e - S_s) + p_s, S_e + S_s, or (S_s - S_e) + p_e.  Written on
    // (x, N) these are (S_n - S_x) + p_x, S_n + S_x and (S_x - S_n) + p_n; the form is uniform
    // over the pass, so it is applied as exact +-1 / 0 coefficients in a chain of fused
    // multiply-adds — the same roundings, in the same order, as the branches of is_turning(),
    // without per-dimension selects.  The verdict (either projection negative) is symmetric in
    // the two projections, so they need not be told apart either.
    struct PairMode {
        double cn, cx, cpx, cpn;  // rho = fma(cpn, p_n, fma(cpx, p_x, fma(cx, S_x, cn * S_n)))
    };
    NB_HD static PairMode pair_mode(int idx_x, int idx_n) {
        const bool n_is_end = idx_x < idx_n;
        const int a = n_is_end ? idx_x : idx_n, b = n_is_end ? idx_n : idx_x;
        const int mode = (a >= 0 && b >= 0) ? 0 : ((b >= 0 && a < 0) ? 1 : 2);
        PairMode m;
        if (mode == 1) {
            m.cn = 1.0; m.cx = 1.0; m.cpx = 0.0; m.cpn = 0.0;
        } else if ((mode == 0) == n_is_end) {  // (S_n - S_x) + p_x
            m.cn = 1.0; m.cx = -1.0; m.cpx = 1.0; m.cpn = 0.0;
        } else {                               // (S_x - S_n) + p_n
            m.cn = -1.0; m.cx = 1.0; m.cpx = 0.0; m.cpn = 1.0;
        }
        return m;
    }
    // vpn = vr * pn (shared with the kinetic energy)
    NB_HD static void turn_terms(const PairMode& m, double px, double sx, double pn, double sn,
                                 double vr, double vpn, double& acc_n, double& acc_x) {
        const double rho = fma(m.cpn, pn, fma(m.cpx, px, fma(m.cx, sx, m.cn * sn)));
        acc_n += rho * vpn;
        acc_x += rho * (vr * px);
    }

    // ------------------------------------------------- bulk-copy staged streaming pass
    // body(k, b) is called once for every double2 index k < D2 owned by this thread, with
    // b[v * T] = srcs[v][k] read from the shared-memory stage (T = threads of the chain).  Thread 0
    // keeps the copy engine kStages chunks ahead (chunk c + kStages is requested at the hand-over
    // barrier of chunk c), so whole stages per chain are in flight regardless of registers.
    // All sources must be in global memory; the pass starts with a proxy fence + barrier
    // because the sources were written with ordinary stores by this CTA.
    NB_HD bool can_stage() const {
#ifdef __CUDA_ARCH__
        if constexpr (stage_smem_bytes<M, G::kThreads>() > 0)
            return (P->stage_loads & 1) && smem_slots == 0 && var == varg;
#endif
        return false;
    }
    // direction of the next streaming pass: opposite to the last leapfrog's (see stage_loads)
    NB_HD bool next_sweep_rev() const {
#ifdef __CUDA_ARCH__
        if constexpr (M::kElementwise && NIT == 0) return (P->stage_loads & 2) ? !sweep_rev : false;
#endif
        return false;
    }
#ifdef __CUDA_ARCH__
    // rev: visit the chunks from the last to the first.  dead_mask / keep_mask: source vectors
    // whose lines are dead once read (L2 evict_first) / re-read by every pass (evict_last).
    template <class F>
    NB_D void staged_pass(const double2* const (&srcs)[kStageVecs], int nv, int D2, bool rev,
                          unsigned dead_mask, unsigned keep_mask, F&& body) {
        constexpr int CH = G::kThreads;              // double2 per vector and chunk
        constexpr int BUF = stage_buf_bytes<CH>();   // bytes per stage
        constexpr int kStages = stage_count<CH>();
        const uint32_t buf0 = nb_smem_u32(stage);
        const uint32_t bar0 = buf0 + kStages * BUF;
        const int nchunk = (D2 + CH - 1) / CH;
        const bool hints = (P->stage_loads & 4) != 0;
        unsigned long long pol_dead = 0, pol_keep = 0;
        if (hints && g.tid == 0) {
            pol_dead = nb_policy_evict_first();
            pol_keep = nb_policy_evict_last();
        }
        if (!hints) dead_mask = keep_mask = 0;
        const int first0 = rev ? (nchunk - 1) * CH : 0, dfirst = rev ? -CH : CH;
        auto issue = [&](int c) {
            const int sidx = c % kStages;
            const int first = first0 + c * dfirst;
            const int n = (D2 - first) < CH ? (D2 - first) : CH;
            const unsigned bytes = (unsigned)n * 16u;
            const uint32_t b = buf0 + sidx * BUF;
            const uint32_t bar = bar0 + 8 * sidx;
            nb_mbar_expect_tx(bar, bytes * (unsigned)nv);
#pragma unroll
            for (int v = 0; v < kStageVecs; ++v) {
                if (v < nv) {
                    if ((dead_mask >> v) & 1u)
                        nb_bulk_g2s_hint(b + v * CH * 16, srcs[v] + first, bytes, bar, pol_dead);
                    else if ((keep_mask >> v) & 1u)
                        nb_bulk_g2s_hint(b + v * CH * 16, srcs[v] + first, bytes, bar, pol_keep);
                    else
                        nb_bulk_g2s(b + v * CH * 16, srcs[v] + first, bytes, bar);
                }
            }
        };
        nb_fence_proxy_async();
        g.sync();
        if (g.tid == 0) {
#pragma unroll
            for (int c = 0; c < kStages; ++c)
                if (c < nchunk) issue(c);
        }
        int k = first0 + g.tid;
        int sidx = 0;
        for (int c = 0; c < nchunk; ++c) {
            nb_mbar_wait(bar0 + 8 * sidx, (stage_phase >> sidx) & 1u);
            stage_phase ^= 1u << sidx;
            if (k < D2) body(k, reinterpret_cast<const double2*>(stage + (size_t)sidx * BUF) + g.tid);
            if (c + kStages < nchunk) {
                // (a consumer-release mbarrier per stage — each warp arrives, thread 0 waits and
                // refills — measured 6 % SLOWER than this CTA barrier: the waiting thread spins
                // inside warp 0 and stalls it; profiles/r2_config4_notes.txt)
                g.sync();  // every thread is done reading this stage
                if (g.tid == 0) issue(c + kStages);
            }
            k += dfirst;
            sidx = sidx + 1 == kStages ? 0 : sidx + 1;
        }
    }
#endif

    // ---------------------------------------------------------------- leapfrog
    // src -> dst (dst is a fresh slot).  Returns 0 ok, 1 divergence.
    // want_l0: also evaluate is_turning(src, dst); verdict in l0_turn.
    // plist[0..np): older states to pair with the new leaf; verdicts in fused_bits.
    NB_HD int leapfrog(int src, int dst, int dir, bool want_l0 = false, const int* plist = nullptr,
                       int np = 0) {
        if constexpr (LR) return leapfrog_lr(src, dst, dir);
        const double eps = (double)dir * step_size;
        const double heps = 0.5 * eps;
        const double* qs = vec(src, VQ);
        const double* ps = vec(src, VP);
        const double* gs = vec(src, VG);
        const double* ss = vec(src, VS);
        double* qd = vec(dst, VQ);
        double* pd = vec(dst, VP);
        double* gd = vec(dst, VG);
        double* sd = vec(dst, VS);
        const int new_idx = sh->idx[src] + dir;
        const bool restart_sum = new_idx == -1;
        const PairMode m_src = pair_mode(new_idx - dir, new_idx);
        const double* Pp[kFusedDim];
        const double* Sp[kFusedDim];
        PairMode Pm[kFusedDim];
#pragma unroll
        for (int c = 0; c < kFusedDim; ++c) {
            const int sl = (c < np) ? plist[c] : src;
            Pp[c] = vec(sl, VP);
            Sp[c] = vec(sl, VS);
            Pm[c] = pair_mode(sh->idx[sl], new_idx);
        }
        fused_bits = 0;
        double lp, kin;
        bool bad;
        if constexpr (M::kElementwise) {
            // one streaming pass: 5 loads + 4 stores per dimension (72 B); logp, kinetic
            // energy and the planned U-turn pairs are reduced together
            constexpr int NA = 5 + 2 * kMaxFused;
            double acc[NA];
#pragma unroll
            for (int c = 0; c < NA; ++c) acc[c] = 0.0;
            // An elementwise gradient is a function of q_i alone: it is recomputed from the
            // source position instead of being stored with every state, which removes one
            // vector read and one vector write per gradient evaluation (72 -> 56 B/dim moved).
            // px / sx: momentum and p_sum of the planned partners at dimension i
            auto elem = [&](int i, double q0, double p0, double vr, double s0, const double (&px)[kFusedDim],
                            const double (&sx)[kFusedDim], double& qn, double& pn, double& sn) {
                double g0, gn;
                (void)M::term(md, i, q0, g0);
                const double ph = p0 + heps * g0;
                qn = q0 + eps * (vr * ph);
                acc[0] += M::term(md, i, qn, gn);
                pn = ph + heps * gn;
                const double vpn = vr * pn;
                acc[1] += pn * vpn;
                sn = restart_sum ? pn : s0 + pn;
                // (a density may promise that a non-finite gradient shows in logp as well)
                if constexpr (!M::kLogpFlagsBadGrad)
                    if (!nb_isfinite(gn)) acc[2] += 1.0;
                if (want_l0) turn_terms(m_src, p0, s0, pn, sn, vr, vpn, acc[3], acc[4]);
#pragma unroll
                for (int c = 0; c < kMaxFused; ++c)
                    if (c < np)
                        turn_terms(Pm[c], px[c], sx[c], pn, sn, vr, vpn, acc[5 + 2 * c], acc[6 + 2 * c]);
            };
            auto elem1 = [&](int i, double& qn, double& pn, double& sn) {  // scalar accesses
                double px[kFusedDim], sx[kFusedDim];
#pragma unroll
                for (int c = 0; c < kFusedDim; ++c) {
                    px[c] = (c < kMaxFused && c < np) ? Pp[c][i] : 0.0;
                    sx[c] = (c < kMaxFused && c < np) ? Sp[c][i] : 0.0;
                }
                elem(i, qs[i], ps[i], var[i], ss[i], px, sx, qn, pn, sn);
            };
            if constexpr (NIT == 0) {
                // 16-byte accesses: two dimensions per thread and iteration (slots are 32-byte
                // aligned, Dp is a multiple of 4); an odd last dimension is handled alone
                const int D2 = D >> 1;
                const double2* qs2 = reinterpret_cast<const double2*>(qs);
                const double2* ps2 = reinterpret_cast<const double2*>(ps);
                const double2* ss2 = reinterpret_cast<const double2*>(ss);
                const double2* vr2 = reinterpret_cast<const double2*>(var);
                double2* qd2 = reinterpret_cast<double2*>(qd);
                double2* pd2 = reinterpret_cast<double2*>(pd);
                double2* sd2 = reinterpret_cast<double2*>(sd);
                bool staged = false;
                const bool rev = next_sweep_rev();
                sweep_rev = rev;
#ifdef __CUDA_ARCH__
                if constexpr (stage_smem_bytes<M, G::kThreads>() > 0) {
                    // Bulk-copy staging: thread 0 asks the copy engine for chunk c+1 of every
                    // source vector while the CTA integrates chunk c out of shared memory, so
                    // the bytes in flight are whole stages (stage_buf_bytes) per chain instead of
                    // what 64 registers per thread can hold.  Needs every source in global memory.
                    staged = can_stage() && D2 > 0;
                    if (staged) {
                        const double2* srcs[kStageVecs];
                        srcs[0] = qs2; srcs[1] = ps2; srcs[2] = vr2; srcs[3] = ss2;
#pragma unroll
                        for (int cc = 0; cc < kMaxFused; ++cc) {
                            srcs[4 + 2 * cc] = reinterpret_cast<const double2*>(Pp[cc]);
                            srcs[5 + 2 * cc] = reinterpret_cast<const double2*>(Sp[cc]);
                        }
                        // the source state and the partners are read once (dead afterwards for
                        // all but tree ends); the mass matrix is read by every pass
                        constexpr int CH = G::kThreads;
                        staged_pass(srcs, 4 + 2 * np, D2, rev, ~4u, 4u, [&](int k, const double2* b) {
                            const double2 q0 = b[0 * CH], p0 = b[1 * CH];
                            const double2 v0 = b[2 * CH], s0 = b[3 * CH];
                            double pxa[kFusedDim], sxa[kFusedDim], pxb[kFusedDim], sxb[kFusedDim];
#pragma unroll
                            for (int cc = 0; cc < kFusedDim; ++cc) {
                                pxa[cc] = sxa[cc] = pxb[cc] = sxb[cc] = 0.0;
                                if (cc < kMaxFused && cc < np) {
                                    const double2 pp = b[(4 + 2 * cc) * CH];
                                    const double2 sp = b[(5 + 2 * cc) * CH];
                                    pxa[cc] = pp.x; pxb[cc] = pp.y;
                                    sxa[cc] = sp.x; sxb[cc] = sp.y;
                                }
                            }
                            double2 qn, pn, sn;
                            elem(2 * k, q0.x, p0.x, v0.x, s0.x, pxa, sxa, qn.x, pn.x, sn.x);
                            elem(2 * k + 1, q0.y, p0.y, v0.y, s0.y, pxb, sxb, qn.y, pn.y, sn.y);
                            qd2[k] = qn;
                            pd2[k] = pn;
                            sd2[k] = sn;
                        });
                    }
                }
#endif
                if (!staged) {
                    const int nth = g.size(), nchunk = (D2 + nth - 1) / nth;
                    for (int c = 0; c < nchunk; ++c) {
                        const int k = (rev ? nchunk - 1 - c : c) * nth + g.tid;
                        if (k >= D2) continue;
                        const double2 q0 = qs2[k], p0 = ps2[k], v0 = vr2[k], s0 = ss2[k];
                        double pxa[kFusedDim], sxa[kFusedDim], pxb[kFusedDim], sxb[kFusedDim];
#pragma unroll
                        for (int cc = 0; cc < kFusedDim; ++cc) {
                            pxa[cc] = sxa[cc] = pxb[cc] = sxb[cc] = 0.0;
                            if (cc < kMaxFused && cc < np) {
                                pxa[cc] = Pp[cc][2 * k]; pxb[cc] = Pp[cc][2 * k + 1];
                                sxa[cc] = Sp[cc][2 * k]; sxb[cc] = Sp[cc][2 * k + 1];
                            }
                        }
                        double2 qn, pn, sn;
                        elem(2 * k, q0.x, p0.x, v0.x, s0.x, pxa, sxa, qn.x, pn.x, sn.x);
                        elem(2 * k + 1, q0.y, p0.y, v0.y, s0.y, pxb, sxb, qn.y, pn.y, sn.y);
                        qd2[k] = qn;
                        pd2[k] = pn;
                        sd2[k] = sn;
                    }
                }
                if ((D & 1) && g.tid == 0) {
                    const int i = D - 1;
                    double qn, pn, sn;
                    elem1(i, qn, pn, sn);
                    qd[i] = qn;
                    pd[i] = pn;
                    sd[i] = sn;
                }
            } else {
                for_dims([&](int i) {
                    double qn, pn, sn;
                    elem1(i, qn, pn, sn);
                    qd[i] = qn;
                    pd[i] = pn;
                    sd[i] = sn;
                });
            }
            if (want_l0 || np > 0) {
                g.reduce(acc);
                l0_turn = (acc[3] < 0.0) | (acc[4] < 0.0);
#pragma unroll
                for (int c = 0; c < kMaxFused; ++c)
                    if (c < np && ((acc[5 + 2 * c] < 0.0) | (acc[6 + 2 * c] < 0.0))) fused_bits |= 1u << c;
            } else {
                double a3[3] = {acc[0], acc[1], acc[2]};
                g.reduce(a3);
                acc[0] = a3[0];
                acc[1] = a3[1];
                acc[2] = a3[2];
            }
            lp = M::finish(md, acc[0], D);
            kin = 0.5 * acc[1];
            bad = acc[2] > 0.0;
        } else {
            // in place on the shared-memory front; the new state is also written out
            // to its pool slot (stores only, nothing waits on them)
            double* fq = front;
            double* fp = front + (size_t)Dp;
            double* fg = front + 2 * (size_t)Dp;
            double* fs = front + 3 * (size_t)Dp;
            if (front_slot != src) {
                for_dims([&](int i) {
                    fq[i] = qs[i];
                    fp[i] = ps[i];
                    fg[i] = gs[i];
                    fs[i] = ss[i];
                });
            }
            front_slot = dst;
            double acc[2] = {0.0, 0.0};
            if constexpr (NIT > 0) {
                // half-step momenta and the mass matrix stay in registers across the
                // density evaluation
                double ph[NIT], vr[NIT];
                NB_PROF_MARK_INIT();
                // every load of the pass before its first store: the front, the mass matrix and the
                // pool slot are reached through pointers the compiler cannot tell apart, so a load
                // placed after a store waits for it (six serialised shared-memory round trips)
                {
                    double q0[NIT];
#pragma unroll
                    for (int it = 0; it < NIT; ++it) {
                        const int i = g.tid + it * G::kThreads;
                        if (it + 1 < NIT || i < D) {
                            vr[it] = var[i];
                            ph[it] = fp[i] + heps * fg[i];
                            q0[it] = fq[i];
                        }
                    }
#pragma unroll
                    for (int it = 0; it < NIT; ++it) {
                        const int i = g.tid + it * G::kThreads;
                        if (it + 1 < NIT || i < D) {
                            const double qn = q0[it] + eps * (vr[it] * ph[it]);
                            fq[i] = qn;
                            qd[i] = qn;
                        }
                    }
                }
                g.sync();
                NB_PROF_MARK(0);
                lp = M::logp_grad(g, md, D, fq, fg, msm);
                g.sync();
                NB_PROF_MARK(1);
                double tacc[2 + 2 * kMaxFused];
#pragma unroll
                for (int c = 0; c < 2 + 2 * kMaxFused; ++c) tacc[c] = 0.0;
                double gnv[NIT], pov[NIT], sov[NIT];
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    const int i = g.tid + it * G::kThreads;
                    if (it + 1 < NIT || i < D) {
                        gnv[it] = fg[i];
                        pov[it] = fp[i];
                        sov[it] = fs[i];
                    }
                }
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    const int i = g.tid + it * G::kThreads;
                    if (it + 1 < NIT || i < D) {
                        const double gn = gnv[it];
                        const double pn = ph[it] + heps * gn;
                        const double vpn = vr[it] * pn;
                        acc[0] += pn * vpn;
                        const double p_old = pov[it], s_old = sov[it];
                        const double sn = restart_sum ? pn : s_old + pn;
                        if (want_l0) turn_terms(m_src, p_old, s_old, pn, sn, vr[it], vpn, tacc[0], tacc[1]);
#pragma unroll
                        for (int c = 0; c < kMaxFused; ++c)
                            if (c < np)
                                turn_terms(Pm[c], Pp[c][i], Sp[c][i], pn, sn, vr[it], vpn, tacc[2 + 2 * c],
                                           tacc[3 + 2 * c]);
                        fp[i] = pn;
                        fs[i] = sn;
                        gd[i] = gn;
                        pd[i] = pn;
                        sd[i] = sn;
                        if (!nb_isfinite(gn)) acc[1] += 1.0;
                    }
                }
                if (want_l0 || np > 0) {
                    double all[4 + 2 * kMaxFused];
                    all[0] = acc[0];
                    all[1] = acc[1];
#pragma unroll
                    for (int c = 0; c < 2 + 2 * kMaxFused; ++c) all[2 + c] = tacc[c];
                    g.reduce(all);
                    acc[0] = all[0];
                    acc[1] = all[1];
                    l0_turn = (all[2] < 0.0) | (all[3] < 0.0);
#pragma unroll
                    for (int c = 0; c < kMaxFused; ++c)
                        if (c < np && ((all[4 + 2 * c] < 0.0) | (all[5 + 2 * c] < 0.0))) fused_bits |= 1u << c;
                } else {
                    g.reduce(acc);
                }
                NB_PROF_MARK(2);
            } else {
                for (int i = g.tid; i < D; i += g.size()) {
                    const double ph = fp[i] + heps * fg[i];
                    fp[i] = ph;
                    const double qn = fq[i] + eps * (var[i] * ph);
                    fq[i] = qn;
                    qd[i] = qn;
                }
                g.sync();
                lp = M::logp_grad(g, md, D, fq, fg, msm);
                g.sync();
                for (int i = g.tid; i < D; i += g.size()) {
                    const double gn = fg[i];
                    const double pn = fp[i] + heps * gn;
                    acc[0] += pn * (var[i] * pn);
                    const double sn = restart_sum ? pn : fs[i] + pn;
                    fp[i] = pn;
                    fs[i] = sn;
                    gd[i] = gn;
                    pd[i] = pn;
                    sd[i] = sn;
                    if (!nb_isfinite(gn)) acc[1] += 1.0;
                }
                g.reduce(acc);
            }
            kin = 0.5 * acc[0];
            bad = acc[1] > 0.0;
        }
        int rc = (bad || !nb_isfinite(lp)) ? 1 : 0;  // recoverable logp error (rc 3/4)
        const double U = -lp;
        const double de = (kin + U) - E0;
        if (rc == 0 && (de > st().max_energy_error || !nb_isfinite(de))) rc = 1;
        if (rc != 0) {
            div_src = src;
            div_dst = dst;
        }
        if (g.tid == 0) {
            sh->idx[dst] = new_idx;
            sh->U[dst] = U;
            sh->K[dst] = kin;
        }
#ifndef NB200_EMUL_DROP_BARRIER  // (tests/emul builds one library without this barrier: the
        g.sync();                // negative control of the lane-schedule race check)
#endif
        last_de = de;
        if (skip_acc) return rc;  // pipeline producer: the consumer keeps the collectors
        acc_count += 1;
        if (rc == 0) {
#ifdef __CUDA_ARCH__
            if (G::kThreads == 32 && defer_acc) {
                if ((unsigned)g.tid == n_parked) de_parked = de;
                if (++n_parked == 32u) flush_acc();
            } else
#endif
            {
                const double w = exp(-de);
                const double a = w < 1.0 ? w : 1.0;
                acc_sum += a;
                acc_sym += 2.0 * a / (1.0 + w);
            }
        }
        last_de = de;
        return rc;
    }

    // add the parked leaves' acceptance statistics to the collectors, in leaf order
    NB_HD void flush_acc() {
#ifdef __CUDA_ARCH__
        if constexpr (G::kThreads == 32) {
            if (n_parked == 0) return;
            const double2 r = nb_flush_parked(de_parked, n_parked, acc_sum, acc_sym);
            acc_sum = r.x;
            acc_sym = r.y;
            n_parked = 0;
        }
#endif
    }

    // ------------------------------------------------ low-rank metric (lowrank.cuh)
    // One leapfrog under M^-1 = S (I + V (L - I) V^T) S.  The velocity needs a contraction over
    // ALL dimensions between the momentum half step and the position step, so the pass cannot be
    // fused like the diagonal one.  The leaf under construction lives in the shared-memory front
    // (q, p, grad, velocity): the source state is read from its pool slot once, the short passes
    // and the two k x D contractions with V (streamed from L2) work on shared memory, and the new
    // state is only STORED to its slot — nothing waits on a global round trip between the passes.
    // The velocity of every state is kept in the slot (VV): the U-turn criterion projects on it.
    NB_HD int leapfrog_lr(int src, int dst, int dir) {
        const double eps = (double)dir * step_size;
        const double heps = 0.5 * eps;
        const double* qs = gvec(src, VQ);
        const double* ps = gvec(src, VP);
        const double* gs = gvec(src, VG);
        const double* ss = gvec(src, VS);
        double* qd = gvec(dst, VQ);
        double* pd = gvec(dst, VP);
        double* gd = gvec(dst, VG);
        double* sd = gvec(dst, VS);
        double* vd = gvec(dst, VV);
        double* fq = front;
        double* fp = front + (size_t)Dp;
        double* fg = front + 2 * (size_t)Dp;
        double* fv = front + 3 * (size_t)Dp;
        front_slot = -1;  // the front is this function's scratch, not a mirror of a slot
        const int new_idx = sh->idx[src] + dir;
        const bool restart_sum = new_idx == -1;
        for (int i = g.tid; i < D; i += g.size()) {
            fp[i] = ps[i] + heps * gs[i];
            fq[i] = qs[i];
        }
        g.sync();
        lr_velocity(g, lr, D, Dp, fp, fv);
        for (int i = g.tid; i < D; i += g.size()) {
            const double qn = fq[i] + eps * fv[i];
            fq[i] = qn;
            qd[i] = qn;
        }
        g.sync();
        double lp;
        double flag[2] = {0.0, 0.0};
        if constexpr (M::kElementwise) {
            for (int i = g.tid; i < D; i += g.size()) {
                double gn;
                flag[1] += M::term(md, i, fq[i], gn);
                fg[i] = gn;
                if (!nb_isfinite(gn)) flag[0] += 1.0;
            }
            g.reduce(flag);
            lp = M::finish(md, flag[1], D);
        } else {
            lp = M::logp_grad(g, md, D, fq, fg, msm);
            g.sync();
            for (int i = g.tid; i < D; i += g.size())
                if (!nb_isfinite(fg[i])) flag[0] += 1.0;
            g.reduce(flag);
        }
        const bool bad = flag[0] > 0.0;
        for (int i = g.tid; i < D; i += g.size()) {
            const double gn = fg[i];
            fp[i] = fp[i] + heps * gn;
            gd[i] = gn;
        }
        g.sync();
        lr_velocity(g, lr, D, Dp, fp, fv);
        double acc[1] = {0.0};
        for (int i = g.tid; i < D; i += g.size()) {
            const double pn = fp[i], vn = fv[i];
            acc[0] += pn * vn;
            pd[i] = pn;
            vd[i] = vn;
            sd[i] = restart_sum ? pn : ss[i] + pn;
        }
        g.reduce(acc);
        const double kin = 0.5 * acc[0];
        int rc = (bad || !nb_isfinite(lp)) ? 1 : 0;
        const double U = -lp;
        const double de = (kin + U) - E0;
        if (rc == 0 && (de > st().max_energy_error || !nb_isfinite(de))) rc = 1;
        if (rc != 0) {
            div_src = src;
            div_dst = dst;
        }
        if (g.tid == 0) {
            sh->idx[dst] = new_idx;
            sh->U[dst] = U;
            sh->K[dst] = kin;
        }
        g.sync();
        last_de = de;
        acc_count += 1;
        if (rc == 0) {
            const double w = exp(-de);
            const double a = w < 1.0 ? w : 1.0;
            acc_sum += a;
            acc_sym += 2.0 * a / (1.0 + w);
        }
        l0_turn = false;
        fused_bits = 0;
        return rc;
    }
    // U-turn criterion on the stored velocities (oracle: is_turning_v)
    NB_HD bool is_turning_lr(int s1, int s2) const {
        int a = sh->idx[s1], b = sh->idx[s2];
        int ss_ = s1, se_ = s2;
        if (!(a < b)) {
            int t = a; a = b; b = t;
            ss_ = s2; se_ = s1;
        }
        const int mode = (a >= 0 && b >= 0) ? 0 : ((b >= 0 && a < 0) ? 1 : 2);
        const double* p_s = gvec(ss_, VP);
        const double* s_s = gvec(ss_, VS);
        const double* v_s = gvec(ss_, VV);
        const double* p_e = gvec(se_, VP);
        const double* s_e = gvec(se_, VS);
        const double* v_e = gvec(se_, VV);
        double acc[2] = {0.0, 0.0};
        for (int i = g.tid; i < D; i += g.size()) {
            double rho;
            if (mode == 0) rho = s_e[i] - s_s[i] + p_s[i];
            else if (mode == 1) rho = s_e[i] + s_s[i];
            else rho = s_s[i] - s_e[i] + p_e[i];
            acc[0] += rho * v_e[i];
            acc[1] += rho * v_s[i];
        }
        g.reduce(acc);
        return (acc[0] < 0.0) | (acc[1] < 0.0);
    }
    // fresh momentum p = M^1/2 z, v = M^-1 p, p_sum = p, K, idx = 0, E0
    NB_HD void init_momentum_lr(int slot, uint32_t purpose, uint32_t rng_draw) {
        double* pd = gvec(slot, VP);
        double* sd = gvec(slot, VS);
        double* vd = gvec(slot, VV);
        const double* tape = nullptr;
        if (purpose == RNG_MOMENTUM && P->z_tape)
            tape = P->z_tape + ((size_t)chain_local * P->n_total + rng_draw) * (size_t)D;
        for (int j = g.tid; 2 * j < D; j += g.size()) {
            double z0, z1;
            const int i0 = 2 * j, i1 = 2 * j + 1;
            if (tape) {
                z0 = tape[i0];
                z1 = i1 < D ? tape[i1] : 0.0;
            } else {
                uint64_t a, b;
                rng_u64x2(st().seed, chain_gid, rng_draw, purpose, (uint32_t)j, a, b);
                rng_normal_pair(a, b, z0, z1);
            }
            pd[i0] = z0;
            if (i1 < D) pd[i1] = z1;
        }
        g.sync();
        lr_momentum(g, lr, D, Dp, pd);
        lr_velocity(g, lr, D, Dp, pd, vd);
        double acc[1] = {0.0};
        for (int i = g.tid; i < D; i += g.size()) {
            const double p = pd[i];
            sd[i] = p;
            acc[0] += p * vd[i];
        }
        g.reduce(acc);
        const double kin = 0.5 * acc[0];
        if (g.tid == 0) {
            sh->idx[slot] = 0;
            sh->K[slot] = kin;
        }
        E0 = kin + sh->U[slot];
        front_slot = -1;
        g.sync();
    }

    // ------------------------------------------------------------------ U-turn
    NB_HD bool is_turning(int s1, int s2) const {
        if constexpr (LR) return is_turning_lr(s1, s2);
        int a = sh->idx[s1], b = sh->idx[s2];
        int ss_ = s1, se_ = s2;
        if (!(a < b)) {
            int t = a; a = b; b = t;
            ss_ = s2; se_ = s1;
        }
        const int mode = (a >= 0 && b >= 0) ? 0 : ((b >= 0 && a < 0) ? 1 : 2);
        double acc[2] = {0.0, 0.0};
        auto body = [&](const double* p_s, const double* sum_s, const double* p_e,
                        const double* sum_e) {
            for_dims([&](int i) {
                const double pse = sum_e[i], pss = sum_s[i], pe = p_e[i], ps = p_s[i];
                double rho;
                if (mode == 0) rho = pse - pss + ps;
                else if (mode == 1) rho = pse + pss;
                else rho = pss - pse + pe;
                const double vr = var[i];
                acc[0] += rho * (vr * pe);
                acc[1] += rho * (vr * ps);
            });
        };
#ifdef __CUDA_ARCH__
        if constexpr (stage_smem_bytes<M, G::kThreads>() > 0) {
            if (can_stage() && D >= 2) {
                // same pass through the bulk-copy stage (two dimensions per thread and chunk)
                const int D2 = D >> 1;
                const double2* srcs[kStageVecs];
                srcs[0] = reinterpret_cast<const double2*>(vec(ss_, VP));
                srcs[1] = reinterpret_cast<const double2*>(vec(ss_, VS));
                srcs[2] = reinterpret_cast<const double2*>(vec(se_, VP));
                srcs[3] = reinterpret_cast<const double2*>(vec(se_, VS));
                srcs[4] = reinterpret_cast<const double2*>(var);
#pragma unroll
                for (int v = 5; v < kStageVecs; ++v) srcs[v] = srcs[0];
                // the three forms of rho as exact +-1 / 0 coefficients (same roundings as `body`)
                const double c_se = mode == 2 ? -1.0 : 1.0, c_ss = mode == 0 ? -1.0 : 1.0;
                const double c_ps = mode == 0 ? 1.0 : 0.0, c_pe = mode == 2 ? 1.0 : 0.0;
                auto term = [&](double ps, double pss, double pe, double pse, double vr) {
                    const double rho = fma(c_pe, pe, fma(c_ps, ps, fma(c_ss, pss, c_se * pse)));
                    acc[0] += rho * (vr * pe);
                    acc[1] += rho * (vr * ps);
                };
                // sweep against the last leapfrog: the newest leaf's tail is still in L2
                const_cast<ChainCtx*>(this)->staged_pass(srcs, 5, D2, next_sweep_rev(), 0u, 16u,
                                                         [&](int, const double2* b) {
                    constexpr int CH = G::kThreads;
                    const double2 ps = b[0 * CH], pss = b[1 * CH];
                    const double2 pe = b[2 * CH], pse = b[3 * CH];
                    const double2 vr = b[4 * CH];
                    term(ps.x, pss.x, pe.x, pse.x, vr.x);
                    term(ps.y, pss.y, pe.y, pse.y, vr.y);
                });
                if ((D & 1) && g.tid == 0) {
                    const int i = D - 1;
                    term(vec(ss_, VP)[i], vec(ss_, VS)[i], vec(se_, VP)[i], vec(se_, VS)[i], var[i]);
                }
                g.reduce(acc);
                return (acc[0] < 0.0) | (acc[1] < 0.0);
            }
        }
#endif
        if constexpr (!M::kElementwise) {
            // the newest leaf is read from the shared-memory front
            const double* f_p = front + (size_t)Dp;
            const double* f_s = front + 3 * (size_t)Dp;
            // One copy of the loop behind selected (generic) pointers: the three
            // address-space-specialised copies it replaces were 3x the code for a saved
            // generic-address decode, and this kernel is bound by instruction fetch
            // (profiles/r1_sweep_radon_codesize.txt: +6 %; an out-of-line single copy for the
            // four call sites of transition() measured no better).
            const double* p_s = ss_ == front_slot ? f_p : vec(ss_, VP);
            const double* s_s = ss_ == front_slot ? f_s : vec(ss_, VS);
            const double* p_e = se_ == front_slot ? f_p : vec(se_, VP);
            const double* s_e = se_ == front_slot ? f_s : vec(se_, VS);
            // the three forms of rho as exact +-1 / 0 coefficients (same roundings as `body`)
            const double c_se = mode == 2 ? -1.0 : 1.0, c_ss = mode == 0 ? -1.0 : 1.0;
            const double c_ps = mode == 0 ? 1.0 : 0.0, c_pe = mode == 2 ? 1.0 : 0.0;
            for_dims([&](int i) {
                const double pse = s_e[i], pss = s_s[i], pe = p_e[i], ps = p_s[i];
                const double rho = fma(c_pe, pe, fma(c_ps, ps, fma(c_ss, pss, c_se * pse)));
                const double vr = var[i];
                acc[0] += rho * (vr * pe);
                acc[1] += rho * (vr * ps);
            });
        } else {
            body(vec(ss_, VP), vec(ss_, VS), vec(se_, VP), vec(se_, VS));
        }
        g.reduce(acc);
        return (acc[0] < 0.0) | (acc[1] < 0.0);
    }

    // ------------------------------------------------------- fresh momentum
    // p = z / sqrt(var), p_sum = p, K, idx = 0, E0 = K + U   (initialize_trajectory)
    NB_HD void init_momentum(int slot, uint32_t purpose, uint32_t rng_draw) {
        if constexpr (LR) {
            init_momentum_lr(slot, purpose, rng_draw);
            return;
        }
        double* pd = vec(slot, VP);
        double* sd = vec(slot, VS);
        if constexpr (!M::kElementwise) {
            // the front takes over (q, grad) of this point; its momentum is written below
            if (front_slot != slot) {
                const double* q = vec(slot, VQ);
                const double* gr = vec(slot, VG);
                double* fq = front;
                double* fg = front + 2 * (size_t)Dp;
                for_dims([&](int i) {
                    fq[i] = q[i];
                    fg[i] = gr[i];
                });
            }
            front_slot = slot;
        }
        double* fp = M::kElementwise ? pd : front + (size_t)Dp;
        double* fs = M::kElementwise ? sd : front + 3 * (size_t)Dp;
        const double* tape = nullptr;
        if (purpose == RNG_MOMENTUM && P->z_tape)
            tape = P->z_tape + ((size_t)chain_local * P->n_total + rng_draw) * (size_t)D;
        double acc[1] = {0.0};
        for (int j = g.tid; 2 * j < D; j += g.size()) {
            double z0, z1;
            const int i0 = 2 * j, i1 = 2 * j + 1;
            if (tape) {
                z0 = tape[i0];
                z1 = i1 < D ? tape[i1] : 0.0;
            } else {
                uint64_t a, b;
                rng_u64x2(st().seed, chain_gid, rng_draw, purpose, (uint32_t)j, a, b);
                rng_normal_pair(a, b, z0, z1);
            }
            {
                const double vr = var[i0];
                const double p = sqrt(1.0 / vr) * z0;
                pd[i0] = p;
                sd[i0] = p;
                if (!M::kElementwise) {
                    fp[i0] = p;
                    fs[i0] = p;
                }
                acc[0] += p * (vr * p);
            }
            if (i1 < D) {
                const double vr = var[i1];
                const double p = sqrt(1.0 / vr) * z1;
                pd[i1] = p;
                sd[i1] = p;
                if (!M::kElementwise) {
                    fp[i1] = p;
                    fs[i1] = p;
                }
                acc[0] += p * (vr * p);
            }
        }
        g.reduce(acc);
        const double kin = 0.5 * acc[0];
        if (g.tid == 0) {
            sh->idx[slot] = 0;
            sh->K[slot] = kin;
        }
        E0 = kin + sh->U[slot];  // U[slot] was written before the previous sync
        g.sync();
    }


    // ------------------------------------------------- two-warp pipeline of one chain
    // The tree bookkeeping of leaf k (U-turn checks against older states, log-sum-exp weights,
    // multinomial draws, acceptance statistics: ~1/3 of a leaf's instructions) does not feed the
    // integration of leaf k + 1, only the decision to STOP.  So a chain is owned by two warps: the
    // PRODUCER integrates leaves back to back (it knows the doubling directions — Philox keyed by
    // draw and depth — and tracks the two trajectory ends itself), the CONSUMER builds the tree
    // from them and ends the draw when it turns / diverges / reaches maxdepth; leaves produced
    // past that point are dropped.  Same arithmetic in the same order as the one-warp kernel:
    // traces are bit-identical.  The critical path per leaf becomes max(integrate, bookkeep)
    // instead of their sum.
#ifdef __CUDA_ARCH__
    NB_D PipeShared* pp() const { return &sh->pipe; }
    NB_D void pipe_init_barriers() {  // one lane, before the chain's warps meet
        PipeShared* ps = pp();
        for (int e = 0; e < kPipeDepth; ++e) {
            nb_mbar_init(&ps->full[e], 1);
            nb_mbar_init(&ps->empty[e], 1);
        }
        nb_mbar_init(&ps->cmd, 1);
        ps->quit = 0;
        ps->stop = 0;
        nb_mbar_init_fence();
    }
    // consumer: hand entry (kgrant % depth) a fresh destination slot
    NB_D void pipe_grant(int slot) {
        PipeShared* ps = pp();
        const unsigned e = kgrant % kPipeDepth;
        if (slot >= 0) reserved |= 1ull << slot;
        if (g.tid == 0) ps->dst[e] = slot;
        __syncwarp();
        if (g.tid == 0) nb_mbar_arrive(nb_smem_u32(&ps->empty[e]));
        kgrant += 1;
    }
    NB_D void pipe_begin_draw(int cur, uint32_t t) {
        PipeShared* ps = pp();
        front_slot = -1;  // the front buffer belongs to the producer during the draw
        reserved = 0;
        if (g.tid == 0) {
            ps->cur = cur;
            ps->draw = t;
            ps->step_size = step_size;
        }
        for (int e = 0; e < kPipeDepth; ++e) pipe_grant(alloc());
        __syncwarp();
        if (g.tid == 0) nb_mbar_arrive(nb_smem_u32(&ps->cmd));
    }
    // consumer: next leaf (blocks until the producer has filled it)
    NB_D int pipe_recv(int& dst) {
        PipeShared* ps = pp();
        const unsigned e = kcons % kPipeDepth;
        {
            NB_PROF_T0();
            nb_mbar_wait(nb_smem_u32(&ps->full[e]), (kcons / kPipeDepth) & 1u);
            NB_PROF_ADD(0);
            NB_PROF_INC(1);
        }
        kcons += 1;
        dst = ps->dst[e];
        const int rc = ps->rc[e];
        const double de = ps->de[e];
        l0_turn = ps->l0[e] != 0;
        reserved &= ~(1ull << dst);
        acc_count += 1;
        if (rc == 0) {
            if (defer_acc) {
                if ((unsigned)g.tid == n_parked) de_parked = de;
                if (++n_parked == 32u) flush_acc();
            } else {
                const double w = exp(-de);
                const double a = w < 1.0 ? w : 1.0;
                acc_sum += a;
                acc_sym += 2.0 * a / (1.0 + w);
            }
        }
        last_de = de;
        return rc;
    }
    // consumer: the draw is over — have the producer skip what is still granted, collect those
    // entries, then post the end-of-draw marker and wait for its acknowledgement
    NB_D void pipe_end_draw() {
        PipeShared* ps = pp();
        NB_PROF_T0();
        if (g.tid == 0) *(volatile int*)&ps->stop = 1;
        while (kcons < kgrant) {
            const unsigned e = kcons % kPipeDepth;
            nb_mbar_wait(nb_smem_u32(&ps->full[e]), (kcons / kPipeDepth) & 1u);
            kcons += 1;
        }
        pipe_grant(-1);
        {
            const unsigned e = kcons % kPipeDepth;
            nb_mbar_wait(nb_smem_u32(&ps->full[e]), (kcons / kPipeDepth) & 1u);
            kcons += 1;
        }
        E0 = ps->E0;
        reserved = 0;
        __syncwarp();
        if (g.tid == 0) *(volatile int*)&ps->stop = 0;
        __syncwarp();
        NB_PROF_ADD(2);
        NB_PROF_INC(3);
    }
    NB_D void pipe_quit() {
        PipeShared* ps = pp();
        if (g.tid == 0) ps->quit = 1;
        __syncwarp();
        if (g.tid == 0) nb_mbar_arrive(nb_smem_u32(&ps->cmd));
    }
    // producer: the whole life of the integrator warp
    NB_D void producer_main() {
        PipeShared* ps = pp();
        skip_acc = true;
        defer_acc = false;
        unsigned k = 0;
        const int maxdepth = (int)st().maxdepth;
        for (;;) {
            {
                NB_PROF_T0();
                nb_mbar_wait(nb_smem_u32(&ps->cmd), ncmd & 1u);
                NB_PROF_ADD(4);
            }
            ncmd += 1;
            if (ps->quit) return;
            const int cur = ps->cur;
            const uint32_t t = ps->draw;
            step_size = ps->step_size;
            front_slot = -1;
            {
                NB_PROF_T0();
                init_momentum(cur, RNG_MOMENTUM, t);
                NB_PROF_ADD(5);
            }
            if (g.tid == 0) ps->E0 = E0;
            int end_f = cur, end_b = cur;
            bool over = false;
            for (int depth = 0; depth < maxdepth && !over; ++depth) {
                uint64_t ra, rb;
                rng_u64x2(st().seed, chain_gid, t, RNG_DIRECTION, (uint32_t)depth, ra, rb);
                const int dir = (ra & 1) ? 1 : -1;
                const bool check = st().check_turning && depth >= (int)st().mindepth;
                int src = dir > 0 ? end_f : end_b;
                const unsigned n_leaf = 1u << depth;
                for (unsigned j = 0; j < n_leaf; ++j) {
                    const unsigned e = k % kPipeDepth;
                    {
                        NB_PROF_T0();
                        nb_mbar_wait(nb_smem_u32(&ps->empty[e]), (k / kPipeDepth) & 1u);
                        NB_PROF_ADD(6);
                    }
                    k += 1;
                    const int dst = ps->dst[e];
                    if (dst >= 0 && !*(volatile int*)&ps->stop) {
                        const bool fuse_l0 = kFuseL0 && check && ((j & 1u) || depth == 0);
                        NB_PROF_T0();
                        const int rc = leapfrog(src, dst, dir, fuse_l0);
                        NB_PROF_ADD(7);
                        NB_PROF_INC(8);
                        if (g.tid == 0) {
                            ps->de[e] = last_de;
                            ps->rc[e] = rc;
                            ps->l0[e] = l0_turn ? 1 : 0;
                        }
                        src = dst;
                    }
                    __syncwarp();
                    if (g.tid == 0) nb_mbar_arrive(nb_smem_u32(&ps->full[e]));
                    if (dst < 0) {
                        over = true;
                        break;
                    }
                }
                if (dir > 0) end_f = src;
                else end_b = src;
            }
            if (!over) {  // every leaf up to maxdepth is out: the end-of-draw marker follows
                for (;;) {
                    const unsigned e = k % kPipeDepth;
                    nb_mbar_wait(nb_smem_u32(&ps->empty[e]), (k / kPipeDepth) & 1u);
                    k += 1;
                    const int dst = ps->dst[e];
                    __syncwarp();
                    if (g.tid == 0) nb_mbar_arrive(nb_smem_u32(&ps->full[e]));
                    if (dst < 0) break;
                }
            }
        }
    }
#endif

    // uniform for merge number n_merge of draw t (same value as drawing it on the spot)
    NB_HD double merge_uniform(uint32_t t) {
#ifdef __CUDA_ARCH__
        if constexpr (G::kThreads == 32) {
            if (n_merge - u_base >= 32u) {
                u_base = n_merge;
                u_parked = nb_refill_merge_uniforms(st().seed, chain_gid, t, u_base);
            }
            return __shfl_sync(0xffffffffu, u_parked, (int)(n_merge - u_base));
        }
#endif
        uint64_t ra, rb;
        rng_u64x2(st().seed, chain_gid, t, RNG_MERGE, n_merge, ra, rb);
        return rng_u01(ra);
    }

    // ------------------------------------------------------------- transition
    template <bool PIPED = false>
    NB_HD int transition(int cur, uint32_t t, SampleInfo& info) {
        draw = t;
        n_merge = 0;
        u_base = 0x80000000u;  // nothing drawn ahead yet
        acc_sum = acc_sym = 0.0;
        acc_count = 0;
        n_parked = 0;
        defer_acc = true;
        mL = mR = mD = cur;
        tL = tR = tD = -1;
        lv_live = 0;
#ifdef __CUDA_ARCH__
        if constexpr (PIPED) pipe_begin_draw(cur, t);
        else
#endif
            init_momentum(cur, RNG_MOMENTUM, t);
        double m_ls = 0.0;
        int depth = 0;
        unsigned spec_bits = 0;  // speculative C verdicts per level (bit 31: main tree)
        info.diverging = 0;
        info.maxdepth_reached = 0;
        bool done = false;
        const int maxdepth = (int)st().maxdepth;
        while (depth < maxdepth && !done) {
            uint64_t ra, rb;
            rng_u64x2(st().seed, chain_gid, t, RNG_DIRECTION, (uint32_t)depth, ra, rb);
            const int dir = (ra & 1) ? 1 : -1;
            const bool check = st().check_turning && depth >= (int)st().mindepth;
            const unsigned n_leaf = 1u << depth;
            int prev = dir > 0 ? mR : mL;
            lv_live = 0;
            tL = tR = tD = -1;
            bool stop = false;  // the new sub-tree is discarded and the transition ends
            for (unsigned j = 0; j < n_leaf && !stop && !done; ++j) {
                int dst = PIPED ? -1 : alloc();
                // ---- plan the U-turn checks that pair the new leaf N with older states.
                // Merging sub-tree s (earlier) with t (later, ending in N) needs three pairs:
                //   A = (far end of s, N),  B = (near end of s, N)  [depth > 0],
                //   C = (far end of s, first leaf of t)             [depth > 0].
                // A and B are evaluated in the leapfrog that creates N (the pair with N's
                // predecessor is free); C was evaluated speculatively in the leapfrog that
                // created t's first leaf and is kept in spec_bits until the merge.
                const bool last = j + 1 == n_leaf;
                const int cascade = nb_ffsll((unsigned long long)(~j)) - 1;  // trailing ones of j
                const bool fuse_l0 = kFuseL0 && check && ((j & 1u) || depth == 0);
                int plist[kFusedDim];
                int np = 0;
                int spec_level = -1;
                if (check && kMaxFused > 0) {
                    if (j == 0) {
                        if (depth > 0) {
                            spec_level = 31;
                            plist[np++] = dir > 0 ? mL : mR;
                        }
                    } else if (!(j & 1u)) {
                        spec_level = nb_ffsll((unsigned long long)j) - 1;
                        plist[np++] = dir > 0 ? sh->lvL[spec_level] : sh->lvR[spec_level];
                    }
                    for (int k = 1; k < cascade && np < kMaxFused; ++k) {
                        plist[np++] = dir > 0 ? sh->lvL[k] : sh->lvR[k];
                        if (np < kMaxFused) plist[np++] = dir > 0 ? sh->lvR[k] : sh->lvL[k];
                    }
                    if (last && depth > 0 && np < kMaxFused) {
                        plist[np++] = dir > 0 ? mL : mR;
                        if (np < kMaxFused) plist[np++] = dir > 0 ? mR : mL;
                    }
                }
                int rc;
#ifdef __CUDA_ARCH__
                if constexpr (PIPED) {
                    rc = pipe_recv(dst);
                    if (rc != 0) {
                        div_src = prev;
                        div_dst = dst;
                    }
                } else
#endif
                {
                    NB_PROF_T0();
                    rc = leapfrog(prev, dst, dir, fuse_l0, plist, np);
                    NB_PROF_ADD(7);
                    NB_PROF_INC(8);
                }
                if (rc != 0) {
                    info.diverging = 1;
                    stop = true;
                    break;
                }
                int fi = 0;  // next fused verdict to consume, in planning order
                if (spec_level >= 0) {
                    const unsigned bit = 1u << spec_level;
                    spec_bits = (fused_bits & 1u) ? (spec_bits | bit) : (spec_bits & ~bit);
                    fi = 1;
                }
                // verdict of the pair (x, N): fused if it was planned, else a separate pass
                auto pair_with_new = [&](int x, int n_slot) -> bool {
                    if (fi < np) return (fused_bits >> fi++) & 1u;
                    return is_turning(x, n_slot);
                };
                tL = tR = tD = dst;
#ifdef __CUDA_ARCH__
                if constexpr (PIPED) pipe_grant(alloc());  // keep the producer kPipeDepth leaves ahead
#endif
                double t_ls = -last_de;
                int k = 0;
                // Merge the new leaf with the parked siblings of equal depth (binary
                // counter) and — once the sub-tree is complete — with the main tree.
                for (;;) {
                    bool with_main;
                    if ((j >> k) & 1u) with_main = false;
                    else if (last && k == depth) with_main = true;
                    else break;
                    const int sL = with_main ? mL : sh->lvL[k];
                    const int sR = with_main ? mR : sh->lvR[k];
                    const int sD = with_main ? mD : sh->lvD[k];
                    const double s_ls = with_main ? m_ls : sh->lvLS[k];
                    bool turn = false;
                    if (check) {
                        const int far_s = dir > 0 ? sL : sR, near_s = dir > 0 ? sR : sL;
                        if constexpr (kMaxFused > 0) {
                            if (k == 0) {
                                // the single pair (s, N); s is N's predecessor
                                turn = fuse_l0 ? l0_turn : is_turning(far_s, dst);
                            } else {
                                const bool a = pair_with_new(far_s, dst);
                                const bool b = pair_with_new(near_s, dst);
                                const bool c = (spec_bits >> (with_main ? 31 : k)) & 1u;
                                turn = a | b | c;
                            }
                        } else {
                            // No fused partners: every pair is a separate pass.  ONE inlined
                            // copy of is_turning() walked by a rolled loop over the (up to
                            // three) pairs — four copies were 420 more instructions in a loop
                            // that is bound by instruction fetch.  The verdict is an OR, so
                            // stopping at the first turning pair changes nothing.
                            int n_pairs = 3;
                            if (k == 0) {
                                n_pairs = fuse_l0 ? 0 : 1;  // (s, N) with s = N's predecessor
                                turn = fuse_l0 && l0_turn;
                            }
                            const int t_first = dir > 0 ? tL : tR;
#pragma unroll 1
                            for (int w = 0; w < n_pairs && !turn; ++w) {
                                NB_PROF_MARK_INIT();
                                turn = is_turning(w == 1 ? near_s : far_s, w == 2 ? t_first : dst);
                                NB_PROF_MARK(4);
                                if (!piped) NB_PROF_INC(5);
                            }
                        }
                    }
                    const double new_ls = nb_logaddexp(s_ls, t_ls);
                    const double u = merge_uniform(t);
                    n_merge += 1;
                    // multinomial pick inside a sub-tree, biased progressive for the main tree
                    const double ref_ls = with_main ? s_ls : new_ls;
                    const bool take_new = t_ls >= ref_ls || u < exp(t_ls - ref_ls);
                    if (with_main) {
                        if (take_new) mD = tD;
                        if (dir > 0) mR = tR;
                        else mL = tL;
                        m_ls = new_ls;
                        depth += 1;
                        tL = tR = tD = -1;
                        if (turn) done = true;
                        break;
                    }
                    if (!take_new) tD = sD;
                    if (dir > 0) tL = sL;
                    else tR = sR;
                    t_ls = new_ls;
                    // level k is consumed: its slots live on only through (tL, tR, tD).  Parked
                    // sub-trees are disjoint sets of leaves, so clearing cannot hit another level.
                    lv_live &= ~((1ull << sL) | (1ull << sR) | (1ull << sD));
                    ++k;
                    if (turn) {
                        stop = true;
                        break;
                    }
                }
                if (stop || tL < 0) break;  // discarded, or merged into the main tree
                // park the finished sub-tree at its level
                g.sync();  // earlier readers of the level arrays are done
                if (g.tid == 0) {
                    sh->lvL[k] = tL;
                    sh->lvR[k] = tR;
                    sh->lvD[k] = tD;
                    sh->lvLS[k] = t_ls;
                }
                g.sync();
                lv_live |= (1ull << tL) | (1ull << tR) | (1ull << tD);
                tL = tR = tD = -1;
                prev = dst;
            }
            if (stop) done = true;
        }
#ifdef __CUDA_ARCH__
        if constexpr (PIPED) pipe_end_draw();
#endif
        flush_acc();
        defer_acc = false;
        if (!done) info.maxdepth_reached = 1;
        info.depth = depth;
        lv_live = 0;
        tL = tR = tD = -1;
        return mD;
    }

    // -------------------------------------------------------- dual averaging
    NB_HD void da_new(double initial_step) {
        da_log_step = log(initial_step);
        da_log_step_adapted = da_log_step;
        da_hbar = 0.0;
        da_mu = log(10.0 * initial_step);
        da_count = 1;
    }
    NB_HD void da_advance(double accept_stat) {
        const double count = (double)da_count;
        const double w = 1.0 / (count + st().da_t0);
        da_hbar = (1.0 - w) * da_hbar + w * (st().target_accept - accept_stat);
        da_log_step = da_mu - da_hbar * sqrt(count) / st().da_gamma;
        const double mk = pow(count, -st().da_k);
        da_log_step_adapted = mk * da_log_step + (1.0 - mk) * da_log_step_adapted;
        da_count += 1;
    }
    // Adam on the log step size (nuts-rs stepsize/adam.rs [recalled]: beta1 0.9, beta2 0.999,
    // eps 1e-8, ascent on accept - target).  State reuses the dual-averaging slots:
    // da_hbar = first moment, da_mu = second moment, da_count = step counter.
    NB_HD void adam_new(double initial_step) {
        da_log_step = log(initial_step);
        da_log_step_adapted = da_log_step;
        da_hbar = 0.0;
        da_mu = 0.0;
        da_count = 0;
    }
    NB_HD void adam_advance(double accept_stat) {
        const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
        const double grad = accept_stat - st().target_accept;
        da_count += 1;
        da_hbar = b1 * da_hbar + (1.0 - b1) * grad;
        da_mu = b2 * da_mu + (1.0 - b2) * grad * grad;
        const double t = (double)da_count;
        const double m_hat = da_hbar / (1.0 - pow(b1, t));
        const double v_hat = da_mu / (1.0 - pow(b2, t));
        da_log_step += st().adam_learning_rate * m_hat / (sqrt(v_hat) + eps);
        da_log_step_adapted = da_log_step;
    }
    NB_HD void step_new(double initial_step) {
        if (st().step_size_method == 1) adam_new(initial_step);
        else da_new(initial_step);
    }
    NB_HD void step_advance(double accept_stat) {
        if (st().step_size_method == 1) adam_advance(accept_stat);
        else da_advance(accept_stat);
    }
    NB_HD double clamp_step(double s) const {
        const double m = st().max_step_size;
        return (m > 0 && s > m) ? m : s;
    }

    // ------------------------------------------------- initial step-size search
    // point: slot holding (q, grad, U).  Its momentum is overwritten.
    NB_HD void step_size_search(int point, uint32_t rng_draw) {
        if (st().step_size_method == 2) {
            step_size = st().fixed_step_size;
            return;
        }
        const double keep_sum = acc_sum, keep_sym = acc_sym;
        const uint32_t keep_count = acc_count;
        mL = mR = mD = point;
        tL = tR = tD = -1;
        lv_live = 0;
        init_momentum(point, RNG_STEP_INIT, rng_draw);
        const int nxt = alloc();
        step_size = st().initial_step;
        int dir = 1;
        int found = 0;  // 0 searching, 1 settled (new dual average), -1 gave up
        // iteration 0 is the forward trial step that decides whether to double or halve
        for (int it = 0; it <= 100 && found == 0; ++it) {
            acc_sum = 0.0;
            acc_count = 0;
            const int rc = leapfrog(point, nxt, dir);
            if (rc != 0) {
                step_size = st().initial_step;
                found = -1;
                break;
            }
            const double accept = acc_sum;
            if (it == 0) {
                dir = accept > st().target_accept ? 1 : -1;
                continue;
            }
            if (dir > 0) {
                if (accept <= st().target_accept || step_size > 1e5) found = 1;
                else step_size *= 2.0;
            } else {
                if (accept >= st().target_accept || step_size < 1e-10) found = 1;
                else step_size /= 2.0;
            }
        }
        if (found == 0) {
            step_size = st().initial_step;
            found = 1;
        }
        if (found == 1) step_new(step_size);
        acc_sum = keep_sum;
        acc_sym = keep_sym;
        acc_count = keep_count;
    }

    // --------------------------------------------------- estimator / mass matrix
    // One fused pass over the dimensions: Welford update of both estimator sets
    // with (q, grad) of `slot` (if add), then the diagonal refresh from the
    // (post-switch) foreground set (if update).  fg_after = set that is the
    // foreground after the optional switch.
    NB_HD void estimator_pass(int slot, bool add, unsigned long long n0, unsigned long long n1,
                              bool update, int fg_after, unsigned long long fg_count) {
        const double* q = vec(slot, VQ);
        const double* gr = vec(slot, VG);
        double* w0 = wf;                       // set 0: mean_q, m2_q, mean_g, m2_g
        double* w1 = wf + 4 * (size_t)Dp;      // set 1
        const bool use_grad = st().use_grad_based_estimate != 0;
        for (int i = g.tid; i < D; i += g.size()) {
            double fq = 0.0, fg = 0.0;  // m2 of draws / grads in the (post-switch) foreground set
            if (add) {
                const double x = q[i];
                double y;
                if constexpr (M::kElementwise) (void)M::term(md, i, x, y);  // gradients are not stored
                else y = gr[i];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    double* w = s ? w1 : w0;
                    const unsigned long long n = s ? n1 : n0;
                    double a, b;
                    if (n == 1) {
                        w[i] = x;
                        w[Dp + i] = 0.0;
                        w[2 * Dp + i] = y;
                        w[3 * Dp + i] = 0.0;
                        a = 0.0;
                        b = 0.0;
                    } else {
                        const double inv = 1.0 / (double)n;
                        double mean = w[i];
                        double diff = x - mean;
                        mean += diff * inv;
                        a = w[Dp + i] + diff * (x - mean);
                        w[i] = mean;
                        w[Dp + i] = a;
                        mean = w[2 * Dp + i];
                        diff = y - mean;
                        mean += diff * inv;
                        b = w[3 * Dp + i] + diff * (y - mean);
                        w[2 * Dp + i] = mean;
                        w[3 * Dp + i] = b;
                    }
                    if (s == fg_after) {
                        fq = a;
                        fg = b;
                    }
                }
            } else if (update) {
                const double* w = fg_after ? w1 : w0;
                fq = w[Dp + i];
                fg = w[3 * Dp + i];
            }
            if (update) {
                double val = use_grad ? sqrt(fq / fg) : fq / (double)fg_count;
                if (nb_isfinite(val) && val != 0.0) {
                    if (val < kVarLower) val = kVarLower;
                    if (val > kVarUpper) val = kVarUpper;
                    var[i] = val;
                    if (varg != var) varg[i] = val;
                }
            }
        }
        g.sync();
    }

    // GlobalStrategy::adapt, after every draw
    NB_HD void adapt(unsigned long long t, int dslot, const SampleInfo& info) {
        last_mean = acc_count ? acc_sum / (double)acc_count : 0.0;
        last_sym = acc_count ? acc_sym / (double)acc_count : 0.0;
        last_n_steps = acc_count;
        const unsigned long long num_tune = st().num_tune;
        if (t >= num_tune) return;
        const bool fixed = st().step_size_method == 2;
        const unsigned long long early_end =
            (unsigned long long)ceil(st().early_window * (double)num_tune);
        const unsigned long long sw =
            (unsigned long long)ceil(st().step_size_window * (double)num_tune);
        const unsigned long long final_window = sw > num_tune ? 0 : num_tune - sw;
        if (t < final_window) {
            const bool is_early = t < early_end;
            const unsigned long long switch_freq =
                is_early ? st().early_mass_matrix_switch_freq : st().mass_matrix_switch_freq;
            const int didx = sh->idx[dslot];
            const bool is_good = info.diverging ? ((didx < 0 ? -didx : didx) > 4) : (didx != 0);
            if constexpr (LR) {
                // low-rank strategy: the estimators are a window (deque) of draws and gradients;
                // the schedule (switch / refresh / first-change step-size search) is the diagonal one
                if (is_good) {
                    const double* q = gvec(dslot, VQ);
                    const double* gr = gvec(dslot, VG);
                    lr_push(g, lr, D, Dp, q, [&](int i) { return gr[i]; });
                }
                const bool could_switch = (unsigned long long)(lr.len - lr.split) >= switch_freq;
                const bool is_late = switch_freq + t > final_window;
                bool force_update = false;
                if (could_switch && !is_late) {
                    lr_switch(lr);
                    force_update = true;
                }
                bool did_change = false;
                if (force_update || (t - last_update >= st().mass_matrix_update_freq))
                    did_change = lr_update(g, lr, D, Dp, st().mass_matrix_gamma, st().mass_matrix_eigval_cutoff);
                if (did_change) last_update = t;
                if (!fixed) step_advance(is_late ? last_sym : last_mean);
                if (did_change && has_initial_mm) {
                    has_initial_mm = 0;
                    step_size_search(dslot, (uint32_t)t);
                } else if (!fixed) {
                    step_size = clamp_step(exp(da_log_step));
                }
                return;
            }
            if (is_good) {
                cnt0 += 1;
                cnt1 += 1;
            }
            const unsigned long long n0 = cnt0, n1 = cnt1;
            const int bg = 1 - fg_sel;
            const bool could_switch = (bg ? cnt1 : cnt0) >= switch_freq;
            const bool is_late = switch_freq + t > final_window;
            bool force_update = false;
            if (could_switch && !is_late) {
                if (fg_sel) cnt1 = 0;  // the old foreground becomes the empty background
                else cnt0 = 0;
                fg_sel = bg;
                force_update = true;
            }
            bool did_change = false;
            bool update = false;
            if (force_update || (t - last_update >= st().mass_matrix_update_freq))
                update = (fg_sel ? cnt1 : cnt0) >= 3;
            if (is_good || update)
                estimator_pass(dslot, is_good, n0, n1, update, fg_sel, fg_sel ? cnt1 : cnt0);
            did_change = update;
            if (did_change) last_update = t;
            if (!fixed) step_advance(is_late ? last_sym : last_mean);
            if (did_change && has_initial_mm) {
                has_initial_mm = 0;
                step_size_search(dslot, (uint32_t)t);
            } else if (!fixed) {
                step_size = clamp_step(exp(da_log_step));
            }
            return;
        }
        if (fixed) return;
        step_advance(last_sym);
        if (t == num_tune - 1) step_size = clamp_step(exp(da_log_step_adapted));
        else step_size = clamp_step(exp(da_log_step));
    }

    // -------------------------------------------------------------- chain init
    // Model::init_position + GlobalStrategy::init.  Returns 0 or NB200_E*.
    NB_HD int init_chain() {
        const int slot = 0;
        double* q = vec(slot, VQ);
        const int tries = P->q0 ? 1 : (st().num_try_init > 0 ? st().num_try_init : 1);
        bool ok = false;
        double lp = 0.0;
        for (int attempt = 0; attempt < tries && !ok; ++attempt) {
            g.sync();
            if (P->q0) {
                const double* src = P->q0 + (size_t)chain_local * D;
                for (int i = g.tid; i < D; i += g.size()) q[i] = src[i];
            } else {
                for (int j = g.tid; 2 * j < D; j += g.size()) {
                    uint64_t a, b;
                    rng_u64x2(st().seed, chain_gid, (uint32_t)attempt, RNG_INIT_POS, (uint32_t)j, a, b);
                    double e0, e1;
                    if (st().init_kind == 1) {
                        rng_normal_pair(a, b, e0, e1);
                    } else {
                        e0 = st().init_radius * (2.0 * rng_u01(a) - 1.0);
                        e1 = st().init_radius * (2.0 * rng_u01(b) - 1.0);
                    }
                    const int i0 = 2 * j, i1 = 2 * j + 1;
                    q[i0] = (P->init_mean ? P->init_mean[i0] : 0.0) + e0;
                    if (i1 < D) q[i1] = (P->init_mean ? P->init_mean[i1] : 0.0) + e1;
                }
            }
            g.sync();
            bool bad;
            lp = eval_logp(slot, bad);
            ok = !bad;
        }
        if (!ok) return NB200_EINIT;
        if (g.tid == 0) {
            sh->U[slot] = -lp;
            sh->idx[slot] = 0;
            sh->K[slot] = 0.0;
        }
        g.sync();
        // mass matrix from |grad| (normalizing_flow.py:1905-1909) and estimator seeds
        const double* gr = vec(slot, VG);
        double* w0 = wf;
        double* w1 = wf + 4 * (size_t)Dp;
        for (int i = g.tid; i < D; i += g.size()) {
            double a = fabs(gr[i]);
            if (a < kVarLower) a = kVarLower;
            if (a > kVarUpper) a = kVarUpper;
            double val = 1.0 / a;
            if (!nb_isfinite(val)) val = 1.0;
            var[i] = val;
            if (varg != var) varg[i] = val;
            const double x = q[i], y = gr[i];
            w0[i] = x; w0[Dp + i] = 0.0; w0[2 * Dp + i] = y; w0[3 * Dp + i] = 0.0;
            w1[i] = x; w1[Dp + i] = 0.0; w1[2 * Dp + i] = y; w1[3 * Dp + i] = 0.0;
        }
        g.sync();
        if constexpr (LR) {
            // low rank [nuts-rs, recalled]: identity metric, the window seeded with the initial point
            for (int i = g.tid; i < D; i += g.size()) lr.stds[i] = 1.0;
            lr.k = 0;
            lr.len = lr.split = lr.head = 0;
            g.sync();
            lr_push(g, lr, D, Dp, q, [&](int i) { return gr[i]; });
        }
        cnt0 = cnt1 = 1;
        fg_sel = 0;
        has_initial_mm = 1;
        last_update = 0;
        total_steps = 0;
        divergences = 0;
        step_new(st().initial_step);
        step_size = st().initial_step;
        acc_sum = acc_sym = 0.0;
        acc_count = 0;
        step_size_search(slot, 0xFFFFFFFFu);
        return 0;
    }

    // ---------------------------------------------------------------- run
    NB_HD void load(const ChainScalars& s) {
        step_size = s.step_size;
        da_log_step = s.da_log_step; da_log_step_adapted = s.da_log_step_adapted;
        da_hbar = s.da_hbar; da_mu = s.da_mu; da_count = s.da_count;
        cnt0 = s.cnt[0]; cnt1 = s.cnt[1]; last_update = s.last_update;
        total_steps = s.total_steps; divergences = s.divergences;
        fg_sel = s.fg_sel; has_initial_mm = s.has_initial_mm;
        sweep_rev = s.sweep_rev != 0;
        if constexpr (LR) {
            lr.k = s.lr_k; lr.len = s.lr_len; lr.split = s.lr_split; lr.head = s.lr_head;
        }
    }
    NB_HD void store(ChainScalars& s, unsigned long long next_draw, int cur, int status) const {
        s.step_size = step_size;
        s.cur_U = sh->U[cur];
        s.da_log_step = da_log_step; s.da_log_step_adapted = da_log_step_adapted;
        s.da_hbar = da_hbar; s.da_mu = da_mu; s.da_count = da_count;
        s.cnt[0] = cnt0; s.cnt[1] = cnt1; s.last_update = last_update;
        s.total_steps = total_steps; s.divergences = divergences;
        s.latest_n_steps = last_n_steps;
        s.fg_sel = fg_sel; s.has_initial_mm = has_initial_mm;
        s.cur_slot = cur;
        s.draw = next_draw;
        s.status = status;
        s.sweep_rev = sweep_rev ? 1 : 0;
        if constexpr (LR) {
            s.lr_k = lr.k; s.lr_len = lr.len; s.lr_split = lr.split; s.lr_head = lr.head;
        }
    }

    // Runs the chain from its persisted state until finished, stopped or the
    // per-launch draw budget is used up.
    NB_HD void run() {
        ChainScalars& sc = P->sc[chain_local];
        int status = sc.status;
        if (status == 2 || status < 0) return;
        int cur;
        unsigned long long t = sc.draw;
        last_n_steps = 0;
        front_slot = -1;
        if (status == 0) {
            const int rc = init_chain();
            g.sync();
            if (rc != 0) {
                if (g.tid == 0) store(sc, 0, 0, rc);
                return;
            }
            cur = 0;
            t = 0;
        } else {
            load(sc);
            cur = sc.cur_slot;
            if (varg != var)
                for (int i = g.tid; i < D; i += g.size()) var[i] = varg[i];
            if (g.tid == 0) {
                sh->U[cur] = sc.cur_U;
                sh->idx[cur] = 0;
                sh->K[cur] = 0.0;
            }
            g.sync();
        }
        const unsigned long long n_total = P->n_total, num_tune = st().num_tune;
        unsigned long long done_here = 0;
        while (t < n_total) {
            if (P->stop_flag) {
                // ONE read of the host's flag per chain and draw: with several warps per chain
                // (a CTA per chain) each warp reading it on its own could see different values
                // around the host's write and part of the chain would enter transition()'s
                // barriers while the rest leaves the loop
                bool stop_now;
                if (G::kThreads <= 32) {
                    stop_now = *P->stop_flag != 0;  // warp-uniform: one load instruction
                } else {
                    if (g.tid == 0) sh->stop = *P->stop_flag;
                    g.sync();
                    stop_now = sh->stop != 0;
                    g.sync();  // everyone has read it before thread 0 may write the next one
                }
                if (stop_now) break;
            }
            if (P->max_draws_per_launch && done_here >= P->max_draws_per_launch) break;
            // step_size_jitter (src/wrapper.rs:393-407; nuts-rs [recalled]): the step used for a
            // draw is the adapted one times a factor uniform in [1 - j, 1 + j]
            const double step_base = step_size;
            if (st().step_size_jitter > 0.0) {
                uint64_t ja, jb;
                rng_u64x2(st().seed, chain_gid, (uint32_t)t, RNG_JITTER, 0u, ja, jb);
                step_size = step_base * (1.0 + st().step_size_jitter * (2.0 * rng_u01(ja) - 1.0));
            }
            const double step_used = step_size;
            const bool keep = st().save_warmup || t >= num_tune;
            const unsigned long long row = st().save_warmup ? t : t - num_tune;
            // trace layout [row][chain][...]: the rows every chain has finished form one
            // contiguous block, so the host streams them out with linear copies
            const size_t row_off = ((size_t)row * P->n_chains + chain_local);
            if (keep && P->mminv) {
                // store_mass_matrix: mass_matrix_inv (diag) | mass_matrix_stds (low rank)
                double* o = P->mminv + row_off * P->gdim;
                const double* src_ = LR ? lr.stds : var;
                for (int i = g.tid; i < (int)P->gdim; i += g.size()) o[i] = src_[i];
            }
            if constexpr (LR) {
                if (keep && P->eigvals) {  // mass_matrix_eigvals, NaN-padded to the maximal rank
                    double* o = P->eigvals + row_off * (size_t)lr.max_rank;
                    for (int j = g.tid; j < lr.max_rank; j += g.size())
                        o[j] = j < lr.k ? lr.vals[j] : nan("");
                }
            }
            SampleInfo info;
            div_src = div_dst = -1;
            int sel;
#ifdef __CUDA_ARCH__
            if (piped) sel = transition<true>(cur, (uint32_t)t, info);
            else
#endif
                sel = transition<false>(cur, (uint32_t)t, info);
            step_size = step_base;
            total_steps += acc_count;
            divergences += info.diverging;
            if (keep && P->divs && info.diverging && div_src >= 0) {
                // DivergenceInfo of the leapfrog that diverged: written before adapt() may reuse
                // the two slots (rows of draws that did not diverge keep the host's NaN fill)
                const int W4 = (int)P->gdim;
                double* o = P->divs + row_off * 4 * (size_t)W4;
                const double* qs_ = vec(div_src, VQ);
                const double* qe_ = vec(div_dst, VQ);
                const double* ps_ = vec(div_src, VP);
                const double* gs_ = vec(div_src, VG);
                for (int i = g.tid; i < W4; i += g.size()) {
                    double gi;
                    if constexpr (M::kElementwise) (void)M::term(md, i, qs_[i], gi);
                    else gi = gs_[i];
                    o[i] = qs_[i];
                    o[W4 + i] = qe_[i];
                    o[2 * W4 + i] = ps_[i];
                    o[3 * W4 + i] = gi;
                }
            }
            // scalars of the selected state, captured before adapt() may reuse the slot
            const double selU = sh->U[sel], selK = sh->K[sel], E0t = E0;
            const int selIdx = sh->idx[sel];
            g.sync();
            adapt(t, sel, info);
            if (keep) {
                const double* q = vec(sel, VQ);
                double* o = P->draws + row_off * P->sdim;
                if (P->expand) M::expand(g, md, D, q, o);  // constrained values + deterministics
                else
                    for (int i = g.tid; i < (int)P->sdim; i += g.size()) o[i] = q[i];
                if (P->grads) {
                    const double* gr = vec(sel, VG);
                    double* og = P->grads + row_off * P->gdim;
                    for (int i = g.tid; i < (int)P->gdim; i += g.size()) {
                        double gi;
                        if constexpr (M::kElementwise) (void)M::term(md, i, q[i], gi);
                        else gi = gr[i];
                        og[i] = gi;
                    }
                }
                if (g.tid == 0) {
                    double* s = P->stats + row_off * NB200_NSTAT;
                    const double U = selU, K = selK;
                    s[NB200_STAT_DEPTH] = info.depth;
                    s[NB200_STAT_MAXDEPTH_REACHED] = info.maxdepth_reached;
                    s[NB200_STAT_INDEX_IN_TRAJECTORY] = (double)selIdx;
                    s[NB200_STAT_LOGP] = -U;
                    s[NB200_STAT_ENERGY] = K + U;
                    s[NB200_STAT_ENERGY_ERROR] = (K + U) - E0t;
                    s[NB200_STAT_DIVERGING] = info.diverging;
                    s[NB200_STAT_STEP_SIZE] = step_used;
                    s[NB200_STAT_STEP_SIZE_BAR] = exp(da_log_step_adapted);
                    s[NB200_STAT_N_STEPS] = (double)last_n_steps;
                    s[NB200_STAT_MEAN_TREE_ACCEPT] = last_mean;
                    s[NB200_STAT_MEAN_TREE_ACCEPT_SYM] = last_sym;
                    s[NB200_STAT_TUNING] = t < num_tune ? 1.0 : 0.0;
                    s[NB200_STAT_DRAW] = (double)t;
                    s[NB200_STAT_CHAIN] = (double)chain_gid;
                    s[NB200_STAT_RESERVED] = 0.0;
                }
            }
            cur = sel;
            t += 1;
            done_here += 1;
            // every 16 draws (and at the end) fence the trace rows written so far and publish
            // the count: the host streams published rows to its buffers while sampling runs
            if ((t & 15ull) == 0 || t == n_total) {
                nb_threadfence_system();  // the reader is a copy engine on another stream
                g.sync();
                if (g.tid == 0) sc.published = t;
            }
            if (g.tid == 0) {  // progress record, polled by the host on a side stream
                sc.draw = t;
                sc.total_steps = total_steps;
                sc.divergences = divergences;
                sc.latest_n_steps = last_n_steps;
                sc.step_size = step_size;
                sc.status = 1;
            }
        }
        g.sync();
        if (g.tid == 0) store(sc, t, cur, t >= n_total ? 2 : 1);
    }
};

}  // namespace nb200
