// launch_impl.cuh — the __global__ kernels and their launchers, instantiated once
// per density in kernels_<model>.cu (separate translation units compile in
// parallel; nb200_api.cu only sees the declarations in launch.hpp).
#pragma once
#include <cuda_runtime.h>

#include "launch.hpp"
#include "nuts_core.cuh"

namespace nb200 {

static inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

// fixed part of a chain's shared memory (scalars, model scratch, reduction scratch)
template <class M, int W>
static size_t chain_smem_fixed(const typename M::Data& md) {
    size_t b = align16(sizeof(ChainShared));
    b += align16(sizeof(double) * (size_t)M::smem_doubles(md, 32 * W));
    if (W > 1) b += align16(sizeof(double) * W * GroupCuda<W>::kMaxRed);
    return b;  // + the front buffer of non-elementwise densities: see stage_bytes()
}

// shared-memory front (q, p, grad, p_sum of the newest leaf) for densities that gather
// across dimensions; elementwise densities stream straight from the pool
template <class M>
__host__ __device__ inline size_t stage_bytes(int Dp) {
    return M::kElementwise ? 0 : ((4 * sizeof(double) * (size_t)Dp + 15) & ~size_t(15));
}

template <class M, int W, int NIT>
__device__ __forceinline__ void setup_ctx(ChainCtx<M, GroupCuda<W>, NIT>& ctx, const KParams<M>& P,
                                          unsigned long long chain, unsigned char* smem_chain) {
    ctx.g.tid = (W == 1) ? (threadIdx.x & 31) : threadIdx.x;
    ctx.P = &P;
    ctx.md = P.mdata;
    ctx.sh = reinterpret_cast<ChainShared*>(smem_chain);
    size_t off = (sizeof(ChainShared) + 15) & ~size_t(15);
    ctx.msm = reinterpret_cast<double*>(smem_chain + off);
    off += (sizeof(double) * (size_t)M::smem_doubles(P.mdata, 32 * W) + 15) & ~size_t(15);
    ctx.g.red = reinterpret_cast<double*>(smem_chain + off);
    if (W > 1) off += (sizeof(double) * W * GroupCuda<W>::kMaxRed + 15) & ~size_t(15);
    ctx.front = reinterpret_cast<double*>(smem_chain + off);
    ctx.front_slot = -1;
    off += stage_bytes<M>(P.Dp);
    double* svar = reinterpret_cast<double*>(smem_chain + off);
    if (P.var_in_smem) off += (sizeof(double) * P.Dp + 15) & ~size_t(15);
    ctx.spool = reinterpret_cast<double*>(smem_chain + off);
    ctx.smem_slots = P.smem_slots;
    ctx.D = P.D;
    ctx.Dp = P.Dp;
    ctx.NS = P.NS;
    ctx.chain_local = chain;
    ctx.chain_gid = (uint32_t)(P.chain_id_offset + chain);
    ctx.pool = P.pool + (size_t)chain * P.NS * 4 * (size_t)P.Dp;
    ctx.varg = P.var + (size_t)chain * P.Dp;
    ctx.var = P.var_in_smem ? svar : ctx.varg;
    ctx.wf = P.welford + (size_t)chain * 8 * (size_t)P.Dp;
    ctx.mL = ctx.mR = ctx.mD = ctx.tL = ctx.tR = ctx.tD = -1;
    ctx.lv_valid = 0;
}

// The sampler: W warps per chain, CPB chains per CTA (CPB > 1 only for W == 1).
// One persistent launch advances every chain through all its draws.
// Streaming regime (W >= 8): cap registers at 64 so that 1024 threads — up to four chains —
// are resident per SM and one chain's reductions / tree bookkeeping overlap another's
// streaming pass.
template <class M, int W, int NIT>
__global__ void __launch_bounds__(W == 1 ? 256 : 32 * W, W >= 8 ? 1024 / (32 * W) : 1)
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
    ChainCtx<M, GroupCuda<W>, NIT> ctx;
    setup_ctx<M, W, NIT>(ctx, P, chain, smem + block_data + (size_t)local * smem_per_chain);
    ctx.md = md;
    ctx.run();
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

template <class M, int W, int NIT>
static cudaError_t launch_one(const KParams<M>& P, size_t smem_per_chain, size_t block_data, int cpb,
                              int grid, int block, cudaStream_t stream) {
    const size_t smem = block_data + smem_per_chain * cpb;
    cudaError_t e = cudaFuncSetAttribute(nuts_kernel<M, W, NIT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(nuts_kernel<M, W, NIT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    nuts_kernel<M, W, NIT><<<grid, block, smem, stream>>>(P, smem_per_chain, block_data);
    return cudaGetLastError();
}

// (W, NIT) combinations with unrolled per-dimension loops; anything else runs NIT = 0
template <class M>
int supported_nit(int W, int nit) {
    if (W == 1) return (nit == 1 || nit == 2 || nit == 3 || nit == 4 || nit == 6 || nit == 8) ? nit : 0;
    if (W == 2) return (nit == 1 || nit == 2 || nit == 3 || nit == 4) ? nit : 0;
    if (W == 4) return (nit == 1 || nit == 2) ? nit : 0;
    return 0;
}

template <class M>
cudaError_t launch_nuts(int W, int NIT, const KParams<M>& P, size_t smem_per_chain,
                        size_t block_data, int cpb, int grid, int block, cudaStream_t stream) {
#define NB_CASE(WW, NN)                                                                       \
    if (W == WW && NIT == NN)                                                                 \
        return launch_one<M, WW, NN>(P, smem_per_chain, block_data, cpb, grid, block, stream);
    NB_CASE(1, 1) NB_CASE(1, 2) NB_CASE(1, 3) NB_CASE(1, 4) NB_CASE(1, 6) NB_CASE(1, 8) NB_CASE(1, 0)
    NB_CASE(2, 1) NB_CASE(2, 2) NB_CASE(2, 3) NB_CASE(2, 4) NB_CASE(2, 0)
    NB_CASE(4, 1) NB_CASE(4, 2) NB_CASE(4, 0)
    NB_CASE(8, 0) NB_CASE(16, 0) NB_CASE(32, 0)
#undef NB_CASE
    return cudaErrorInvalidValue;
}

template <class M>
cudaError_t launch_component(int W, const KParams<M>& P, int mode, const double* scal, double* out,
                             size_t smem, unsigned n) {
#define NB_CASE(WW)                                                                          \
    if (W == WW) {                                                                           \
        cudaError_t e = cudaFuncSetAttribute(component_kernel<M, WW>,                        \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                             (int)smem);                                     \
        if (e != cudaSuccess) return e;                                                      \
        component_kernel<M, WW><<<n, 32 * WW, smem>>>(P, mode, scal, out);                   \
        return cudaGetLastError();                                                           \
    }
    NB_CASE(1) NB_CASE(2) NB_CASE(4) NB_CASE(8) NB_CASE(16) NB_CASE(32)
#undef NB_CASE
    return cudaErrorInvalidValue;
}

template <class M>
size_t smem_fixed(int W, const typename M::Data& md, int Dp) {
    const size_t st = stage_bytes<M>(Dp);
    switch (W) {
    case 1: return chain_smem_fixed<M, 1>(md) + st;
    case 2: return chain_smem_fixed<M, 2>(md) + st;
    case 4: return chain_smem_fixed<M, 4>(md) + st;
    case 8: return chain_smem_fixed<M, 8>(md) + st;
    case 16: return chain_smem_fixed<M, 16>(md) + st;
    default: return chain_smem_fixed<M, 32>(md) + st;
    }
}

template <class M>
size_t model_block_data_bytes(const typename M::Data& md) {
    if constexpr (M::kHasBlockData) return (M::block_data_bytes(md) + 15) & ~size_t(15);
    else return 0;
}

#define NB200_INSTANTIATE_MODEL(M)                                                              \
    template int supported_nit<M>(int, int);                                                    \
    template cudaError_t launch_nuts<M>(int, int, const KParams<M>&, size_t, size_t, int, int,  \
                                        int, cudaStream_t);                                          \
    template cudaError_t launch_component<M>(int, const KParams<M>&, int, const double*,        \
                                             double*, size_t, unsigned);                        \
    template size_t smem_fixed<M>(int, const M::Data&, int);                                    \
    template size_t model_block_data_bytes<M>(const M::Data&);

}  // namespace nb200
