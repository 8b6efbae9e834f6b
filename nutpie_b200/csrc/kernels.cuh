// kernels.cuh — the __global__ kernels of the engine.  Device-only code: compiled by nvcc
// into kernels_<model>.cu for the built-in densities and by NVRTC at run time for
// NB200_MODEL_CUSTOM (kernels_custom.cu amalgamates this header with the user's source).
#pragma once
#ifdef NB200_PIPE_PROFILE
#include <cstdio>
#endif
#include "nuts_core.cuh"

namespace nb200 {

// shared-memory front (q, p, grad, p_sum of the newest leaf) for densities that gather
// across dimensions; elementwise densities stream straight from the pool
template <class M>
__host__ __device__ inline size_t stage_bytes(int Dp, bool low_rank = false) {
    // (the low-rank engine builds every leaf on the front, whatever the density)
    return (M::kElementwise && !low_rank) ? 0 : ((4 * sizeof(double) * (size_t)Dp + 15) & ~size_t(15));
}

template <class M, int W, int NIT, class G = GroupCuda<W>, bool LR = false>
__device__ __forceinline__ void setup_ctx(ChainCtx<M, G, NIT, LR>& ctx, const KParams<M>& P,
                                          unsigned long long chain, unsigned char* smem_chain) {
    if constexpr (G::kThreads < 32) {  // sub-warp group: L aligned lanes of the warp
        const int lane = threadIdx.x & 31;
        ctx.g.tid = lane % G::kThreads;
        ctx.g.mask = ((1u << G::kThreads) - 1u) << (lane - ctx.g.tid);
    } else {
        ctx.g.tid = (W == 1) ? (threadIdx.x & 31) : threadIdx.x;
    }
    ctx.P = &P;
    ctx.md = P.mdata;
    ctx.sh = reinterpret_cast<ChainShared*>(smem_chain);
    size_t off = (sizeof(ChainShared) + 15) & ~size_t(15);
    ctx.msm = reinterpret_cast<double*>(smem_chain + off);
    off += (sizeof(double) * (size_t)M::smem_doubles(P.mdata, G::kThreads) + 15) & ~size_t(15);
    ctx.g.red = reinterpret_cast<double*>(smem_chain + off);
    if (W > 1) off += (sizeof(double) * W * G::kMaxRed + 15) & ~size_t(15);
    ctx.stage = smem_chain + off;
    ctx.stage_phase = 0;
    ctx.sweep_rev = false;
    if constexpr (stage_smem_bytes<M, 32 * W>() > 0) {
        if (ctx.g.tid == 0) {
            for (int st = 0; st < stage_count<32 * W>(); ++st)
                nb_mbar_init(ctx.stage + stage_count<32 * W>() * stage_buf_bytes<32 * W>() + 8 * st, 1);
            nb_mbar_init_fence();
        }
        __syncthreads();
        off += stage_smem_bytes<M, 32 * W>();
    }
    ctx.front = reinterpret_cast<double*>(smem_chain + off);
    ctx.front_slot = -1;
    off += stage_bytes<M>(P.Dp, LR);
    double* svar = reinterpret_cast<double*>(smem_chain + off);
    if (P.var_in_smem) off += (sizeof(double) * P.Dp + 15) & ~size_t(15);
    ctx.spool = reinterpret_cast<double*>(smem_chain + off);
    ctx.smem_slots = P.smem_slots;
    ctx.D = P.D;
    ctx.Dp = P.Dp;
    ctx.NS = P.NS;
    ctx.chain_local = chain;
    ctx.chain_gid = (uint32_t)(P.chain_id_offset + chain);
    ctx.pool = P.pool + (size_t)chain * P.NS * (LR ? 5 : 4) * (size_t)P.Dp;
    if constexpr (LR) {
        const size_t Dp = (size_t)P.Dp, R = (size_t)P.lr_max_rank;
        ctx.lr.stds = P.lr_stds + chain * Dp;
        ctx.lr.vals = P.lr_vals + chain * R;
        ctx.lr.vecs = P.lr_vecs + chain * R * Dp;
        ctx.lr.coef = P.lr_coef + chain * R;
        ctx.lr.win = P.lr_win + chain * (size_t)P.lr_cap * 2 * Dp;
        ctx.lr.matL = P.lr_mat + chain * 2 * (size_t)P.D * Dp;
        ctx.lr.matW = ctx.lr.matL + (size_t)P.D * Dp;
        ctx.lr.cols = P.lr_cols + chain * 6 * Dp;
        ctx.lr.cap = P.lr_cap;
        ctx.lr.max_rank = P.lr_max_rank;
        ctx.lr.k = ctx.lr.len = ctx.lr.split = ctx.lr.head = 0;
    }
    ctx.varg = P.var + (size_t)chain * P.Dp;
    ctx.var = P.var_in_smem ? svar : ctx.varg;
    ctx.wf = P.welford + (size_t)chain * 8 * (size_t)P.Dp;
    ctx.mL = ctx.mR = ctx.mD = ctx.tL = ctx.tR = ctx.tD = -1;
    ctx.lv_live = 0;
    ctx.n_parked = 0;
    ctx.defer_acc = false;
    if constexpr (M::kNeedsChain) {  // host plug-in: mailbox index + request sequence number
        ctx.md.chain = chain;
        if (ctx.g.tid == 0) M::init_chain(ctx.md, ctx.msm);
        ctx.g.sync();
    }
}

// The sampler: W warps per chain, CPB chains per CTA (CPB > 1 only for W == 1).
// One persistent launch advances every chain through all its draws.
// Streaming regime: four chains resident per SM, so that one chain's reductions / tree
// bookkeeping overlap another's streaming pass — 256 threads per chain with registers capped at
// 64, or 128 threads per chain with 128 registers (bulk-copy staging keeps the bytes in flight
// independent of the thread count, and half as many warps repeat the chain's scalar work).
template <class M, int W, int NIT>
constexpr int kernel_min_blocks() {
    if (W >= 8) return 1024 / (32 * W);
    if (W == 4 && M::kElementwise && NIT == 0) return 4;
    return 1;
}
template <class M, int W, int NIT, bool LR = false>
__global__ void __launch_bounds__(W == 1 ? 256 : 32 * W, kernel_min_blocks<M, W, NIT>())
    nuts_kernel(const __grid_constant__ KParams<M> P, size_t smem_per_chain, size_t block_data) {
    extern __shared__ __align__(16) unsigned char smem[];
    typename M::Data md = P.mdata;
    if constexpr (M::kHasBlockData) {
        if (block_data > 0) {  // CTA-wide copy of the density's constant tables
            M::load_block_data(md, smem, threadIdx.x, blockDim.x);
            __syncthreads();
        }
    }
    const int local = (W == 1) ? (threadIdx.x >> 5) : 0;
    const int cpb = (W == 1) ? (blockDim.x >> 5) : 1;
    const unsigned long long chain = (unsigned long long)blockIdx.x * cpb + local;
    if (chain >= P.n_chains) return;
    ChainCtx<M, GroupCuda<W>, NIT, LR> ctx;
    setup_ctx<M, W, NIT, GroupCuda<W>, LR>(ctx, P, chain, smem + block_data + (size_t)local * smem_per_chain);
    if constexpr (M::kHasBlockData) ctx.md = md;  // tables staged in shared memory
#ifdef NB200_PIPE_PROFILE
    const long long t_begin = clock64();
#endif
    ctx.run();
#ifdef NB200_PIPE_PROFILE
    if (ctx.g.tid == 0 && (chain % 257) == 0)
        printf("chain %llu one-warp total %lld | leapfrog %lld (n %lld) = pre %lld + density %lld + post %lld | "
               "is_turning passes %lld (n %lld)\n", chain, clock64() - t_begin, ctx.prof[7], ctx.prof[8],
               ctx.prof[0], ctx.prof[1], ctx.prof[2], ctx.prof[4], ctx.prof[5]);
    if constexpr (M::kHasBlockData) {
        if (blockIdx.x == 0 && threadIdx.x == 0)
            printf("  density phases (chain 0, cumulative over launches): setup %lld walk %lld scan %lld counties %lld "
                   "reduce %lld tail %lld\n", g_density_prof[0], g_density_prof[1], g_density_prof[2],
                   g_density_prof[3], g_density_prof[4], g_density_prof[5]);
    }
#endif
}

// Sub-warp geometry: L lanes per chain, 32 / L chains per warp, `wpb` warps per CTA (GroupSub).
template <class M, int L, int NIT>
__global__ void __launch_bounds__(256, 1)
    nuts_kernel_sub(const __grid_constant__ KParams<M> P, size_t smem_per_chain) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int CPW = 32 / L;  // chains per warp
    const int local = (threadIdx.x >> 5) * CPW + (threadIdx.x & 31) / L;
    const int cpb = (blockDim.x >> 5) * CPW;
    const unsigned long long chain = (unsigned long long)blockIdx.x * cpb + local;
    if (chain >= P.n_chains) return;
    ChainCtx<M, GroupSub<L>, NIT> ctx;
    setup_ctx<M, 1, NIT, GroupSub<L>>(ctx, P, chain, smem + (size_t)local * smem_per_chain);
    ctx.run();
}

// Two warps per chain (W = 1 density geometry, see ChainCtx::producer_main).  Warps are dealt to
// the SM's four schedulers round-robin (warp % 4): warps 4g, 4g+1 are the PRODUCERS of chains 2g,
// 2g+1 and warps 4g+2, 4g+3 their CONSUMERS, so two schedulers run nothing but the integrator
// loop and two nothing but the tree loop — each scheduler's instruction cache sees one loop
// (mixing the roles on a scheduler made instruction fetch the bottleneck:
// profiles/r2_pipeline_notes.txt).  512 threads = up to 8 chains per CTA, 128 registers.
template <class M, int NIT>
__global__ void __launch_bounds__(512, 1)
    nuts_kernel_piped(const __grid_constant__ KParams<M> P, size_t smem_per_chain, size_t block_data,
                      int cpb) {
    extern __shared__ __align__(16) unsigned char smem[];
    typename M::Data md = P.mdata;
    if constexpr (M::kHasBlockData) {
        if (block_data > 0) {
            M::load_block_data(md, smem, threadIdx.x, blockDim.x);
            __syncthreads();
        }
    }
    const int warp = threadIdx.x >> 5;
    const int local = ((warp >> 2) << 1) | (warp & 1);
    const int role = ((warp >> 1) & 1) ^ 1;  // 1 = producer (integrator), 0 = consumer (tree)
    const unsigned long long chain = (unsigned long long)blockIdx.x * cpb + local;
    if (local >= cpb || chain >= P.n_chains) return;
    ChainCtx<M, GroupCuda<1>, NIT> ctx;
    setup_ctx<M, 1, NIT>(ctx, P, chain, smem + block_data + (size_t)local * smem_per_chain);
    if constexpr (M::kHasBlockData) ctx.md = md;
    ctx.piped = true;
    if (role == 0 && ctx.g.tid == 0) ctx.pipe_init_barriers();
    asm volatile("bar.sync %0, 64;" ::"r"(1 + local) : "memory");  // the chain's two warps
#ifdef NB200_PIPE_PROFILE
    const long long t_begin = clock64();
#endif
    if (role == 0) {
        ctx.run();
        ctx.pipe_quit();
    } else {
        ctx.producer_main();
    }
#ifdef NB200_PIPE_PROFILE
    if (ctx.g.tid == 0 && (chain % 257) == 0)
        printf("chain %llu role %d total %lld | C: wait_full %lld (n %lld) end_draw %lld (n %lld) | "
               "P: wait_cmd %lld momentum %lld wait_empty %lld leapfrog %lld (n %lld)\n",
               chain, role, clock64() - t_begin, ctx.prof[0], ctx.prof[1], ctx.prof[2], ctx.prof[3],
               ctx.prof[4], ctx.prof[5], ctx.prof[6], ctx.prof[7], ctx.prof[8]);
#endif
}

// Component kernel: mode 0 = density at q (slot 0); mode 1 = one leapfrog
// slot 0 -> slot 1 with per-state eps/dir/idx.  scal: [n][4] = eps, dir, idx, unused;
// out_scal: [n][4] = logp, kinetic, rc, unused.
template <class M, int W>
__global__ void __launch_bounds__(32 * W)
    component_kernel(const __grid_constant__ KParams<M> P, int mode, const double* scal,
                     double* out_scal) {
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned long long chain = blockIdx.x;
    if (chain >= P.n_chains) return;
    ChainCtx<M, GroupCuda<W>, 0> ctx;
    setup_ctx<M, W, 0>(ctx, P, chain, smem);
    ctx.acc_sum = ctx.acc_sym = 0.0;
    ctx.acc_count = 0;
    if (mode == 0) {
        bool bad;
        const double lp = ctx.eval_logp(0, bad);
        if (ctx.g.tid == 0) {
            out_scal[chain * 4 + 0] = lp;
            out_scal[chain * 4 + 1] = 0.0;
            out_scal[chain * 4 + 2] = bad ? (isfinite(lp) ? 3.0 : 4.0) : 0.0;
        }
    } else {
        const double eps = scal[chain * 4 + 0];
        const int dir = scal[chain * 4 + 1] > 0 ? 1 : -1;
        if (ctx.g.tid == 0) {
            ctx.sh->idx[0] = (int)scal[chain * 4 + 2];
            ctx.sh->U[0] = 0.0;
            ctx.sh->K[0] = 0.0;
        }
        ctx.g.sync();
        ctx.step_size = eps;
        ctx.E0 = 0.0;
        const int rc = ctx.leapfrog(0, 1, dir);
        if constexpr (M::kElementwise) {
            // elementwise gradients are recomputed on use, not stored with the state: materialise
            // it for the caller of this component entry point
            const double* qn = ctx.vec(1, VQ);
            double* gn = ctx.vec(1, VG);
            for (int i = ctx.g.tid; i < ctx.D; i += ctx.g.size()) {
                double gi;
                (void)M::term(ctx.md, i, qn[i], gi);
                gn[i] = gi;
            }
        }
        if (ctx.g.tid == 0) {
            out_scal[chain * 4 + 0] = -ctx.sh->U[1];
            out_scal[chain * 4 + 1] = ctx.sh->K[1];
            out_scal[chain * 4 + 2] = (double)rc;
        }
    }
}

}  // namespace nb200
