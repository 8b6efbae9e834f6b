// models.cuh — device log-densities (logp + gradient on the unconstrained space).
//
// These replace `CpuLogpFunc::logp` (src/pymc.rs:197-215, src/stan.rs:454-463)
// for the BASELINE.json configs.  The reference obtains the density as a HOST
// function pointer compiled by numba (python/nutpie/compile_pymc.py:970-1006);
// a device engine needs device code, so each density is hand-written here and
// has a host twin in the reference plug-in ABI under oracle/models.c.
//
// Interface (all members static, `G` = thread group owning the chain):
//   kElementwise = true : logp = finish(sum_i term(i, q_i)), g_i local.  The
//       leapfrog fuses these into its single streaming pass (72 B per
//       dimension per gradient evaluation).
//   kElementwise = false: logp_grad(grp, data, D, q, g, sm) is called by all
//       threads of the group after q is visible; it writes g and returns the
//       (group-uniform) logp.
#pragma once
#include "group.cuh"
#include "portable.cuh"

namespace nb200 {

struct RadonObs;
NB_HD RadonObs nb_ldg_obs(const RadonObs* p, bool in_smem);
template <class T>
NB_HD T nb_ld_tab(const T* p, bool in_smem) {
    return in_smem ? *p : nb_ldg(p);
}

#define NB_LOG_2PI 1.8378770664093454835606594728112
#define NB_HALF_LOG_2_OVER_PI (-0.22579135264472743236309761494744)

// logp = -1/2 sum ((x - mu)/sigma)^2 : Stan `x ~ normal(mu, 1)` (README.md:148-163,
// tests/test_stan.py:16-24) and the D = 10 000 bandwidth-bound config 4.
struct NormalModel {
    static constexpr bool kElementwise = true;
    static constexpr bool kNeedsChain = false;
    static constexpr bool kRuntimeLoopsOnly = false;
    static constexpr bool kPipelined = false;  // two-warp producer / consumer kernels exist
    static constexpr bool kSubWarp = true;     // sub-warp (4/8/16 lanes per chain) kernels exist
    static constexpr bool kHasBlockData = false;
    // g_i = -(q_i - mu) / var is non-finite only if q_i - mu is, and then so is the term
    // (q_i - mu)^2 of logp: the leapfrog needs no separate per-dimension gradient check
    static constexpr bool kLogpFlagsBadGrad = true;
    struct Data {
        double mu, inv_var;
    };
    NB_HD static int smem_doubles(const Data&, int) { return 0; }
    NB_HD static double term(const Data& d, int, double q, double& g) {
        double r = q - d.mu;
        g = -r * d.inv_var;
        return r * r;
    }
    NB_HD static double finish(const Data& d, double acc, int) { return -0.5 * acc * d.inv_var; }
    // expand_vector (src/pymc.rs:217-286): no transforms, no deterministics
    NB_HD static int expanded_dim(int D) { return D; }
    template <class G>
    NB_HD static void expand(const G& grp, const Data&, int D, const double* q, double* out) {
        for (int i = grp.tid; i < D; i += grp.size()) out[i] = q[i];
    }
};

// Neal's funnel (docs/sample-stats.qmd:19-21; 9 parameters in BASELINE.json)
struct FunnelModel {
    static constexpr bool kElementwise = false;
    static constexpr bool kNeedsChain = false;
    static constexpr bool kRuntimeLoopsOnly = false;
    static constexpr bool kPipelined = false;  // two-warp producer / consumer kernels exist
    static constexpr bool kSubWarp = true;     // sub-warp (4/8/16 lanes per chain) kernels exist
    static constexpr bool kHasBlockData = false;
    struct Data {
        int unused;
    };
    NB_HD static int smem_doubles(const Data&, int) { return 0; }
    template <class G>
    NB_HD static double logp_grad(const G& grp, const Data&, int D, const double* q, double* g,
                                  double*) {
        const double v = q[0];
        const double e = exp(-2.0 * v);
        double acc[1] = {0.0};
        for (int i = grp.tid; i < D; i += grp.size()) {
            if (i >= 1) {
                double x = q[i];
                acc[0] += x * x;
                g[i] = -x * e;
            }
        }
        grp.reduce(acc);
        const double n = (double)(D - 1);
        if (grp.tid == 0) g[0] = -v + acc[0] * e - n;
        return -0.5 * v * v - 0.5 * acc[0] * e - n * v;
    }
    NB_HD static int expanded_dim(int D) { return D; }
    template <class G>
    NB_HD static void expand(const G& grp, const Data&, int D, const double* q, double* out) {
        for (int i = grp.tid; i < D; i += grp.size()) out[i] = q[i];
    }
};

// Hierarchical radon model (README.md:53-88, plain-Normal raw effects as in
// notebooks/pytensor_logp.md:57-88), D = 2J + 5, parameter order = PyMC value
// variable order: intercept, county_raw[J], log county_sd, floor_effect,
// county_floor_raw[J], log county_floor_sd, log sigma.
//
// Data layout (built on the host for the group size T, radon_layout.hpp): the
// observations are sorted by (county, floor) and cut into T contiguous ranges,
// one per thread, stored transposed ([step][thread]) as 16-byte records
// {y, meta} so a warp's loads coalesce into one LDG.128 per observation.  Inside
// a range, a GROUP is a maximal run of equal (county, floor): all its
// observations share the linear predictor mu[2*county + floor], which every
// thread reads from a shared table (meta carries its byte offset).  The walk keeps a
// running PREFIX sum of residuals over the thread's range and stores it to the group's
// private slot when the end-of-group bit is set.  An exclusive scan of the threads' range
// totals (warp shuffles) turns those into GLOBAL prefixes over the sorted observations, and
// the sum of a (county, floor) pair is the difference of two of them: the prefix where its
// last piece ends minus the prefix where the previous non-empty pair ends — one 16-byte
// table record and eight shared-memory loads per county, however the ranges cut the pairs
// (the first version summed up to kmax piece differences per pair in a rolled loop: 10 % of
// a leapfrog's instructions).  No atomics, no data-dependent branches, bitwise reproducible
// (determinism contract, tests/test_stan.py:67-101).
struct RadonObs {
    double y;
    int32_t meta;  // byte offset of mu[2*county+floor] | 1 if last observation of its group
    int32_t pad;
};

NB_HD RadonObs nb_ldg_obs(const RadonObs* p, bool in_smem) {
#ifdef __CUDA_ARCH__
    int4 v;
    if (in_smem) {
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"((unsigned)__cvta_generic_to_shared(p)));
    } else {
        v = __ldg(reinterpret_cast<const int4*>(p));
    }
    RadonObs r;
    r.y = __hiloint2double(v.y, v.x);
    r.meta = v.z;
    r.pad = 0;
    return r;
#else
    return *p;
#endif
}

// three exponentials: on the device lanes 0..2 of each warp evaluate one each and the
// results are broadcast (same instruction count as ONE exp for the warp); bit-identical on
// every lane because each value is computed exactly once
template <class G>
NB_HD void nb_exp3(const G&, double a, double b, double c, double& ea, double& eb, double& ec) {
#ifdef __CUDA_ARCH__
    const int lane = threadIdx.x & 31;
    const double x = lane == 0 ? a : (lane == 1 ? b : c);
    const double e = exp(x);
    ea = __shfl_sync(0xffffffffu, e, 0);
    eb = __shfl_sync(0xffffffffu, e, 1);
    ec = __shfl_sync(0xffffffffu, e, 2);
#else
    ea = exp(a);
    eb = exp(b);
    ec = exp(c);
#endif
}

#if defined(NB200_PIPE_PROFILE) && defined(__CUDACC__)
// cycle profile of the density's phases (chain 0 of the launch only; printed by nuts_kernel)
__device__ long long g_density_prof[8];
#endif
#if defined(NB200_PIPE_PROFILE) && defined(__CUDA_ARCH__)
#define NB_DP_INIT() long long dp_mark_ = clock64(); const bool dp_on_ = blockIdx.x == 0 && threadIdx.x == 0
#define NB_DP_MARK(i)                                             \
    do {                                                          \
        const long long now_ = clock64();                         \
        if (dp_on_) g_density_prof[i] += now_ - dp_mark_;         \
        dp_mark_ = now_;                                          \
    } while (0)
#else
#define NB_DP_INIT()
#define NB_DP_MARK(i)
#endif
struct RadonModel {
    static constexpr bool kElementwise = false;
    static constexpr bool kNeedsChain = false;
    static constexpr bool kRuntimeLoopsOnly = false;
    static constexpr bool kPipelined = true;   // two-warp producer / consumer kernels exist
    static constexpr bool kSubWarp = false;
    // The observation records and group tables are the same for every chain: a CTA
    // that hosts several chains copies them into shared memory once (launch_impl.cuh).
    static constexpr bool kHasBlockData = true;
    struct Data {
        int J, N, n_steps, G, kmax;  // counties, observations, steps/thread, group slots, pieces per pair
        int T, in_smem;              // group size the layout was built for; tables live in shared memory
        const RadonObs* obs;         // [n_steps][T]; padding: y = 0, meta -> mu[2J] (= 0), no end bit
        const int32_t* group_base;   // [T]   first group slot of each thread (one spare slot each)
        const uint32_t* group_list;  // [J][4]: per county {slot | prev_slot << 16, thread |
                                     // prev_thread << 16} for floor 0, then floor 1: where the
                                     // pair's last piece ends / where the previous pair's does
    };
    // expand_vector for the radon model (src/pymc.rs:217-286; producer
    // python/nutpie/compile_pymc.py:816-861): value variables on the constrained scale
    // (the three log-transformed scales exponentiated) followed by the Deterministics
    // county_effect = county_raw * county_sd and county_floor_effect likewise: 4J + 5 values,
    // same order as oracle_expand_radon.
    NB_HD static int expanded_dim(int D) { return 2 * D - 5; }
    template <class G>
    NB_HD static void expand(const G& grp, const Data& d, int D, const double* q, double* out) {
        const int J = d.J;
        const double sd_a = exp(q[J + 1]), sd_b = exp(q[2 * J + 3]);
        for (int i = grp.tid; i < D; i += grp.size()) {
            double v = q[i];
            if (i == J + 1) v = sd_a;
            else if (i == 2 * J + 3) v = sd_b;
            else if (i == 2 * J + 4) v = exp(v);
            out[i] = v;
        }
        for (int c = grp.tid; c < J; c += grp.size()) {
            out[D + c] = q[1 + c] * sd_a;
            out[D + J + c] = q[J + 3 + c] * sd_b;
        }
    }
    NB_HD static size_t block_data_bytes(const Data& d) {
        size_t b = sizeof(RadonObs) * (size_t)d.n_steps * d.T;
        b += (sizeof(uint32_t) * (size_t)4 * d.J + 15) & ~size_t(15);
        b += (sizeof(int32_t) * (size_t)d.T + 15) & ~size_t(15);
        return b;
    }
#ifdef __CUDACC__
    // all threads of the CTA copy the tables; the caller synchronises
    __device__ static void load_block_data(Data& d, unsigned char* dst, int tid, int nthreads) {
        const int n16 = d.n_steps * d.T;  // 16-byte records
        int4* o = reinterpret_cast<int4*>(dst);
        const int4* src = reinterpret_cast<const int4*>(d.obs);
        for (int i = tid; i < n16; i += nthreads) o[i] = __ldg(src + i);
        size_t off = sizeof(RadonObs) * (size_t)n16;
        uint32_t* gl = reinterpret_cast<uint32_t*>(dst + off);
        for (int i = tid; i < 4 * d.J; i += nthreads) gl[i] = __ldg(d.group_list + i);
        off += (sizeof(uint32_t) * (size_t)4 * d.J + 15) & ~size_t(15);
        int32_t* gb = reinterpret_cast<int32_t*>(dst + off);
        for (int i = tid; i < d.T; i += nthreads) gb[i] = __ldg(d.group_base + i);
        d.obs = reinterpret_cast<const RadonObs*>(dst);
        d.group_list = gl;
        d.group_base = gb;
        d.in_smem = 1;
    }
#endif
    // mu[2J+1] (last = 0 for padding) + gsum[G+1] (last = 0) + thread offsets[T+1] (last = 0)
    NB_HD static int smem_doubles(const Data& d, int T) { return 2 * d.J + 1 + d.G + 1 + T + 1; }

    template <class G>
    NB_HD static double logp_grad(const G& grp, const Data& d, int, const double* q, double* g,
                                  double* sm) {
        const int J = d.J, T = grp.size();
        const double intercept = q[0];
        const double log_sd_a = q[J + 1];
        const double floor_eff = q[J + 2];
        const double log_sd_b = q[2 * J + 3];
        const double log_sigma = q[2 * J + 4];
        double sd_a, sd_b, sigma;
        NB_DP_INIT();
        nb_exp3(grp, log_sd_a, log_sd_b, log_sigma, sd_a, sd_b, sigma);
        const double inv_sigma = 1.0 / sigma;
        const double inv_s2 = inv_sigma * inv_sigma;
        double* mu = sm;                // [2J+1] linear predictor per (county, floor)
        double* gsum = sm + 2 * J + 1;  // [G+1]  running prefix of residuals at each group end
        double* toff = gsum + d.G + 1;  // [T+1]  sum of the residuals of all earlier threads
#pragma unroll 1
        for (int c = grp.tid; c < J; c += T) {
            const double a = intercept + q[1 + c] * sd_a;
            mu[2 * c] = a;
            mu[2 * c + 1] = a + (floor_eff + q[J + 3 + c] * sd_b);
        }
        if (grp.tid == 0) {
            mu[2 * J] = 0.0;
            gsum[d.G] = 0.0;
            toff[T] = 0.0;
        }
        grp.sync();
        NB_DP_MARK(0);  // exp, 1 / sigma, linear predictors
        double ss0 = 0.0, ss1 = 0.0, ss2 = 0.0, ss3 = 0.0;
        double range_sum;
        {
            char* kp = reinterpret_cast<char*>(gsum + nb_ld_tab(d.group_base + grp.tid, d.in_smem));
            const char* mub = reinterpret_cast<const char*>(mu);
            const RadonObs* ob = d.obs + grp.tid;
            double pre = 0.0;  // running PREFIX sum of residuals over the thread's whole range
            auto step = [&](const RadonObs& rc, double& ss) {
                const int mt = rc.meta;
                const double r = rc.y - *reinterpret_cast<const double*>(mub + (mt & ~7));
                ss += r * r;
                pre += r;
                if (mt & 1) {  // last observation of its group: publish the prefix sum
                    *reinterpret_cast<double*>(kp) = pre;
                    kp += 8;
                }
            };
            // four records are loaded before they are consumed; the 0..3 left over go one by one
            int j0 = 0;
            // `mu` is read-only during the walk, but it shares the scratch with the prefix slots the
            // walk writes: the four predictors of a group are loaded BEFORE its first (conditional)
            // store, or each load would wait behind the store in front of it
            auto step_m = [&](const RadonObs& rc, double m, double& ss) {
                const double r = rc.y - m;
                ss += r * r;
                pre += r;
                if (rc.meta & 1) {  // last observation of its group: publish the prefix sum
                    *reinterpret_cast<double*>(kp) = pre;
                    kp += 8;
                }
            };
            // (a depth-one software pipeline across groups — next group's records and predictors
            // requested before this group's stores — measured 7 % SLOWER: more code in a loop that
            // is bound by instruction fetch, and the first register spills; gpurun_out/r2_s29)
            for (; j0 + 4 <= d.n_steps; j0 += 4) {
                RadonObs rec[4];
                double m[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) rec[u] = nb_ldg_obs(ob + (size_t)(j0 + u) * T, d.in_smem);
#pragma unroll
                for (int u = 0; u < 4; ++u) m[u] = *reinterpret_cast<const double*>(mub + (rec[u].meta & ~7));
                step_m(rec[0], m[0], ss0);
                step_m(rec[1], m[1], ss1);
                step_m(rec[2], m[2], ss2);
                step_m(rec[3], m[3], ss3);
            }
#pragma unroll 1  // at most three steps: not worth 100 instructions of unrolled remainder
            for (; j0 < d.n_steps; ++j0) step(nb_ldg_obs(ob + (size_t)j0 * T, d.in_smem), ss0);
            range_sum = pre;
        }
        NB_DP_MARK(1);  // observation walk
        // global prefix = thread-local prefix + what all earlier threads summed
        toff[grp.tid] = grp.exclusive_scan(range_sum);
        grp.sync();
        NB_DP_MARK(2);  // scan
        double acc[7] = {(ss0 + ss1) + (ss2 + ss3), 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        // kept rolled (as is the piece loop inside): the kernel is bound by instruction fetch, and
        // unrolling these two cost 660 SASS instructions for nothing (+3 % rolled, measured)
#pragma unroll 1
        for (int c = grp.tid; c < J; c += T) {
            // one 16-byte record per county: the sum of a (county, floor) pair is the difference
            // of two global prefixes (end of its last piece, end of the previous pair's)
            uint32_t e0, t0, e1, t1;
#ifdef __CUDA_ARCH__
            const uint4 rec4 = d.in_smem ? *reinterpret_cast<const uint4*>(d.group_list + 4 * c)
                                         : __ldg(reinterpret_cast<const uint4*>(d.group_list + 4 * c));
            e0 = rec4.x; t0 = rec4.y; e1 = rec4.z; t1 = rec4.w;
#else
            e0 = d.group_list[4 * c]; t0 = d.group_list[4 * c + 1];
            e1 = d.group_list[4 * c + 2]; t1 = d.group_list[4 * c + 3];
#endif
            const double S0 = (gsum[e0 & 0xFFFFu] + toff[t0 & 0xFFFFu]) - (gsum[e0 >> 16] + toff[t0 >> 16]);
            const double S1 = (gsum[e1 & 0xFFFFu] + toff[t1 & 0xFFFFu]) - (gsum[e1 >> 16] + toff[t1 >> 16]);
            const double E = (S0 + S1) * inv_s2;  // sum over the county of d logp / d mu_i
            const double F = S1 * inv_s2;         // same, floor = 1 observations only
            const double ra = q[1 + c], rb = q[J + 3 + c];
            acc[1] += ra * sd_a * E;
            acc[2] += rb * sd_b * F;
            acc[3] += ra * ra;
            acc[4] += rb * rb;
            acc[5] += E;
            acc[6] += F;
            g[1 + c] = sd_a * E - ra;
            g[J + 3 + c] = sd_b * F - rb;
        }
        NB_DP_MARK(3);  // county loop
        grp.reduce(acc);
        NB_DP_MARK(4);  // reduction
        const double ssn = acc[0] * inv_s2;
        // constants: -1/2 (N + 2J + 2) log 2pi - log 10 - log 2 + 3/2 log(2/pi) - log 1.5
        const double kConst = -0.5 * NB_LOG_2PI * (double)(d.N + 2 * J + 2) - 2.3025850929940456840 -
                              0.69314718055994530942 + 3.0 * NB_HALF_LOG_2_OVER_PI -
                              0.40546510810816438198;
        double logp = kConst - 0.5 * ssn - d.N * log_sigma;
        logp += -0.5 * (acc[3] + acc[4]);
        logp += -0.005 * intercept * intercept - 0.125 * floor_eff * floor_eff;
        logp += -0.5 * sd_a * sd_a + log_sd_a;
        logp += -0.5 * sd_b * sd_b + log_sd_b;
        logp += -0.5 * sigma * sigma * (1.0 / 2.25) + log_sigma;
        if (grp.tid == 0) {
            g[0] = acc[5] - 0.01 * intercept;
            g[J + 1] = acc[1] - sd_a * sd_a + 1.0;
            g[J + 2] = acc[6] - 0.25 * floor_eff;
            g[2 * J + 3] = acc[2] - sd_b * sd_b + 1.0;
            g[2 * J + 4] = ssn - d.N - sigma * sigma * (1.0 / 2.25) + 1.0;
        }
        NB_DP_MARK(5);  // scalar tail
        return logp;
    }
};

// ---- the reference's HOST plug-in (NB200_MODEL_HOST, include/nutpie_b200.h) ----
// `int logp(size_t dim, const double* x, double* grad, double* logp, const void* user_data)`
// (RawLogpFunc, src/pymc.rs:23-29) is a host function: the chain posts its position to a
// per-chain mailbox in MAPPED PINNED host memory, rings a doorbell, and waits for the answer a
// host thread writes back (nb200_api.cu: HostService).  Everything else — integrator, tree,
// adaptation — stays in the persistent kernel, so any existing numba cfunc / LogpFunc samples
// through the same engine.  Slow by construction (one PCIe round trip + a host call per
// gradient): the fallback of SURVEY.md §8f-3, not the product path.
struct HostModel {
    static constexpr bool kElementwise = false;
    static constexpr bool kHasBlockData = false;
    static constexpr bool kRuntimeLoopsOnly = true;  // only the NIT = 0 kernels are instantiated
    static constexpr bool kPipelined = false;
    static constexpr bool kSubWarp = false;
    struct Data {
        double* qbox;        // [n_chains][Dp]  device -> host
        double* gbox;        // [n_chains][Dp]  host -> device
        double* lpbox;       // [n_chains]
        int* rcbox;          // [n_chains]      return code of the host function
        unsigned* req;       // [n_chains]      doorbell: sequence number of the posted request
        unsigned* resp;      // [n_chains]      sequence number of the last answered request
        const volatile int* stop;  // the sampler's stop flag (abort must not hang on a dead host)
        unsigned long long chain;  // set per chain by the kernel (kNeedsChain)
        int Dp;
    };
    static constexpr bool kNeedsChain = true;
    NB_HD static int smem_doubles(const Data&, int) { return 2; }  // [0] sequence number, [1] status
    // Sequence numbers continue from the last doorbell this chain ever rang (a launch that was
    // stopped while a request was in flight leaves req ahead of resp: the next request must
    // still be a NEW number for the host thread to notice it).
    NB_HD static void init_chain(const Data& d, double* sm) {
#ifdef __CUDA_ARCH__
        unsigned last;
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(last) : "l"(d.req + d.chain) : "memory");
        reinterpret_cast<unsigned*>(sm)[0] = last;
#else
        (void)d; (void)sm;
#endif
    }
    template <class G>
    NB_HD static double logp_grad(const G& grp, const Data& d, int D, const double* q, double* g,
                                  double* sm) {
#ifdef __CUDA_ARCH__
        const size_t base = (size_t)d.chain * (size_t)d.Dp;
        for (int i = grp.tid; i < D; i += grp.size()) d.qbox[base + i] = q[i];
        __threadfence_system();
        grp.sync();
        unsigned* seqp = reinterpret_cast<unsigned*>(sm);
        if (grp.tid == 0) {
            const unsigned seq = seqp[0] + 1u;
            seqp[0] = seq;
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(d.req + d.chain), "r"(seq) : "memory");
            unsigned got, ns = 2000u;
            int status = 0;
            unsigned long long t0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            for (;;) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(d.resp + d.chain) : "memory");
                if (got == seq) break;
                if (d.stop && *d.stop) { status = 1; break; }  // aborted while waiting
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 60000000000ull) { status = 2; break; }  // 60 s: the host is gone
                __nanosleep(ns);
                if (ns < 32000u) ns *= 2u;
            }
            seqp[1] = (unsigned)status;
        }
        grp.sync();
        const int status = (int)seqp[1];
        double lp = __longlong_as_double(0x7ff8000000000000ll);
        int rc = 1;
        if (status == 0) {
            for (int i = grp.tid; i < D; i += grp.size()) g[i] = __ldcv(d.gbox + base + i);
            lp = __ldcv(d.lpbox + d.chain);
            rc = __ldcv(d.rcbox + d.chain);
        } else {
            for (int i = grp.tid; i < D; i += grp.size()) g[i] = 0.0;
        }
        // rc > 0: recoverable (src/pymc.rs:178) -> the trajectory diverges here; rc < 0 is
        // fatal: the host service has already raised the stop flag and recorded the code
        return rc != 0 ? __longlong_as_double(0x7ff8000000000000ll) : lp;
#else
        (void)grp; (void)d; (void)D; (void)q; (void)g; (void)sm;
        return 0.0;  // device only
#endif
    }
    // expand_vector runs on the host as well (nb200_host_expand_rows): draws stay unconstrained
    NB_HD static int expanded_dim(int D) { return D; }
    template <class G>
    NB_HD static void expand(const G& grp, const Data&, int D, const double* q, double* out) {
        for (int i = grp.tid; i < D; i += grp.size()) out[i] = q[i];
    }
};

}  // namespace nb200

// ---- run-time compiled densities (NB200_MODEL_CUSTOM, include/nutpie_b200.h) ----
// The user's CUDA source defines nb200_user_logp(); kernels_custom.cu appends it to an
// amalgamation of these headers and compiles the sampler kernel with NVRTC for the (threads
// per chain, dimensions per thread) geometry the host picked, so a custom density runs inside
// the same persistent kernel as the built-in ones — the device analogue of the reference's
// numba-compiled LogpFunc (src/pymc.rs:50-62, compile_pymc.py:970-1006).
#ifdef NB200_RTC
struct nb200_group {
    int tid, nthreads;
    double* scratch;  // nb200_model_desc::n_user_scratch doubles of shared memory, private to the chain
    const void* impl_;
    // bit-reproducible sum over the chain's threads; every thread must call it
    __device__ double sum(double x) const {
        double a[1] = {x};
        static_cast<const nb200::GroupCuda<NB200_RTC_W>*>(impl_)->reduce(a);
        return a[0];
    }
    __device__ void sync() const { static_cast<const nb200::GroupCuda<NB200_RTC_W>*>(impl_)->sync(); }
};
__device__ int nb200_user_logp(const nb200_group& grp, int dim, const double* q, double* grad,
                               double* logp_partial, const double* data);
#endif

namespace nb200 {

struct CustomModel {
    static constexpr bool kElementwise = false;
    static constexpr bool kNeedsChain = false;
    static constexpr bool kRuntimeLoopsOnly = false;
    static constexpr bool kPipelined = false;  // two-warp producer / consumer kernels exist
    static constexpr bool kSubWarp = false;    // sub-warp (4/8/16 lanes per chain) kernels exist
    static constexpr bool kHasBlockData = false;
    struct Data {
        const double* data;  // device copy of nb200_model_desc::user_data
        int n_data;
        int program;         // host-side registry slot of the compiled source
        int n_scratch;       // doubles of per-chain shared scratch the density asked for
    };
    NB_HD static int smem_doubles(const Data& d, int) { return d.n_scratch; }
    template <class G>
    NB_HD static double logp_grad(const G& grp, const Data& d, int D, const double* q, double* g,
                                  double* sm) {
#ifdef NB200_RTC
        nb200_group ug{grp.tid, grp.size(), sm, &grp};
        double part = 0.0;
        const int rc = nb200_user_logp(ug, D, q, g, &part, d.data);
        double acc[2] = {part, rc != 0 ? 1.0 : 0.0};
        grp.reduce(acc);
        // rc > 0 is the reference's recoverable error (src/pymc.rs:178): the engine treats a
        // non-finite logp as exactly that (divergent transition / rejected initial point)
        return acc[1] > 0.0 ? __longlong_as_double(0x7ff8000000000000ll) : acc[0];
#else
        (void)grp; (void)d; (void)D; (void)q; (void)g; (void)sm;
        return 0.0;  // only ever instantiated by NVRTC
#endif
    }
    NB_HD static int expanded_dim(int D) { return D; }
    template <class G>
    NB_HD static void expand(const G& grp, const Data&, int D, const double* q, double* out) {
        for (int i = grp.tid; i < D; i += grp.size()) out[i] = q[i];
    }
};

}  // namespace nb200
