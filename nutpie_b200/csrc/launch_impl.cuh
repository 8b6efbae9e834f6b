// launch_impl.cuh — the __global__ kernels and their launchers, instantiated once
// per density in kernels_<model>.cu (separate translation units compile in
// parallel; nb200_api.cu only sees the declarations in launch.hpp).
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "launch.hpp"

namespace nb200 {

static inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

// fixed part of a chain's shared memory (scalars, model scratch, reduction scratch)
template <class M, int W>
static size_t chain_smem_fixed(const typename M::Data& md) {
    size_t b = align16(sizeof(ChainShared));
    b += align16(sizeof(double) * (size_t)M::smem_doubles(md, 32 * W));
    if (W > 1) b += align16(sizeof(double) * W * GroupCuda<W>::kMaxRed);
    b += (size_t)stage_smem_bytes<M, 32 * W>();  // bulk-copy staging of the streaming leapfrog
    return b;  // + the front buffer of non-elementwise densities: see stage_bytes()
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

// low-rank adaptation (lowrank.cuh): the run-time-loop engine with five vectors per slot
template <class M, int W>
static cudaError_t launch_one_lr(const KParams<M>& P, size_t smem_per_chain, size_t block_data, int cpb,
                                 int grid, int block, cudaStream_t stream) {
    const size_t smem = block_data + smem_per_chain * cpb;
    cudaError_t e = cudaFuncSetAttribute(nuts_kernel<M, W, 0, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    nuts_kernel<M, W, 0, true><<<grid, block, smem, stream>>>(P, smem_per_chain, block_data);
    return cudaGetLastError();
}
template <class M>
cudaError_t launch_nuts_lr(int W, const KParams<M>& P, size_t smem_per_chain, size_t block_data,
                           int cpb, int grid, int block, cudaStream_t stream) {
    if (W == 1) return launch_one_lr<M, 1>(P, smem_per_chain, block_data, cpb, grid, block, stream);
    if (W == 4) return launch_one_lr<M, 4>(P, smem_per_chain, block_data, cpb, grid, block, stream);
    return cudaErrorInvalidValue;
}

template <class M, int NIT>
static cudaError_t launch_one_piped(const KParams<M>& P, size_t smem_per_chain, size_t block_data,
                                    int cpb, int grid, cudaStream_t stream) {
    const size_t smem = block_data + smem_per_chain * cpb;
    cudaError_t e = cudaFuncSetAttribute(nuts_kernel_piped<M, NIT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(nuts_kernel_piped<M, NIT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    // warps come in groups of four (two producers, two consumers): see nuts_kernel_piped
    nuts_kernel_piped<M, NIT><<<grid, 128 * ((cpb + 1) / 2), smem, stream>>>(P, smem_per_chain, block_data, cpb);
    return cudaGetLastError();
}

// ---- sub-warp geometry (GroupSub<L>): L lanes per chain, densities with kSubWarp
template <class M, int L, int NIT>
static cudaError_t launch_one_sub(const KParams<M>& P, size_t smem_per_chain, int wpb, int grid,
                                  cudaStream_t stream) {
    const size_t smem = smem_per_chain * (size_t)wpb * (32 / L);
    cudaError_t e = cudaFuncSetAttribute(nuts_kernel_sub<M, L, NIT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    nuts_kernel_sub<M, L, NIT><<<grid, 32 * wpb, smem, stream>>>(P, smem_per_chain);
    return cudaGetLastError();
}

// (lanes per chain, unrolled dimensions per lane) pairs that exist; 0 = none
template <class M>
int sub_warp_lanes(int D, int forced_lanes) {
    if constexpr (M::kSubWarp) {
        if (forced_lanes == 4) return D <= 12 ? 4 : 0;  // up to three dimensions per lane
        if (forced_lanes == 8 || forced_lanes == 16) return D <= 2 * forced_lanes ? forced_lanes : 0;
        if (forced_lanes != 0) return 0;
        if (D <= 4) return 4;
        if (D <= 16) return 8;
        return 0;
    } else {
        (void)D; (void)forced_lanes;
        return 0;
    }
}

template <class M>
cudaError_t launch_nuts_sub(int L, int NIT, const KParams<M>& P, size_t smem_per_chain, int wpb,
                            int grid, cudaStream_t stream) {
    if constexpr (M::kSubWarp) {
#define NB_CASE(LL, NN) \
    if (L == LL && NIT == NN) return launch_one_sub<M, LL, NN>(P, smem_per_chain, wpb, grid, stream);
        NB_CASE(4, 1) NB_CASE(4, 2) NB_CASE(4, 3) NB_CASE(8, 1) NB_CASE(8, 2) NB_CASE(16, 1) NB_CASE(16, 2)
#undef NB_CASE
    }
    return cudaErrorInvalidValue;
}

// two-warp pipeline kernels exist for these unrolled trip counts (densities with kPipelined)
template <class M>
bool supports_pipeline(int NIT) {
    if constexpr (M::kPipelined) return NIT == 2 || NIT == 3 || NIT == 4 || NIT == 6 || NIT == 8;
    else return false;
}

template <class M>
cudaError_t launch_nuts_piped(int NIT, const KParams<M>& P, size_t smem_per_chain, size_t block_data,
                              int cpb, int grid, cudaStream_t stream) {
    if constexpr (M::kPipelined) {
#define NB_CASE(NN) \
    if (NIT == NN) return launch_one_piped<M, NN>(P, smem_per_chain, block_data, cpb, grid, stream);
        NB_CASE(2) NB_CASE(3) NB_CASE(4) NB_CASE(6) NB_CASE(8)
#undef NB_CASE
    }
    return cudaErrorInvalidValue;
}

// (W, NIT) combinations with unrolled per-dimension loops; anything else runs NIT = 0
template <class M>
int supported_nit(int W, int nit) {
    if constexpr (M::kRuntimeLoopsOnly) return 0;
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
    NB_CASE(1, 0) NB_CASE(2, 0) NB_CASE(4, 0) NB_CASE(8, 0) NB_CASE(16, 0) NB_CASE(32, 0)
    if constexpr (!M::kRuntimeLoopsOnly) {
    NB_CASE(1, 1) NB_CASE(1, 2) NB_CASE(1, 3) NB_CASE(1, 4) NB_CASE(1, 6) NB_CASE(1, 8)
    NB_CASE(2, 1) NB_CASE(2, 2) NB_CASE(2, 3) NB_CASE(2, 4)
    NB_CASE(4, 1) NB_CASE(4, 2)
    }
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
    template bool supports_pipeline<M>(int);                                                    \
    template int sub_warp_lanes<M>(int, int);                                                   \
    template cudaError_t launch_nuts_sub<M>(int, int, const KParams<M>&, size_t, int, int,      \
                                            cudaStream_t);                                      \
    template cudaError_t launch_nuts_piped<M>(int, const KParams<M>&, size_t, size_t, int, int, \
                                              cudaStream_t);                                    \
    template cudaError_t launch_nuts_lr<M>(int, const KParams<M>&, size_t, size_t, int, int, int,  \
                                           cudaStream_t);                                          \
    template size_t model_block_data_bytes<M>(const M::Data&);

}  // namespace nb200
