// launch.hpp — declarations of the per-density kernel launchers (defined in
// launch_impl.cuh, instantiated in kernels_<model>.cu).
#pragma once
#include <cuda_runtime.h>

#include "nuts_core.cuh"

namespace nb200 {

template <class M>
int supported_nit(int W, int nit);
template <class M>
cudaError_t launch_nuts(int W, int NIT, const KParams<M>& P, size_t smem_per_chain,
                        size_t block_data, int cpb, int grid, int block, cudaStream_t stream);
template <class M>
int sub_warp_lanes(int D, int forced_lanes);
template <class M>
cudaError_t launch_nuts_sub(int L, int NIT, const KParams<M>& P, size_t smem_per_chain, int wpb,
                            int grid, cudaStream_t stream);
template <class M>
bool supports_pipeline(int NIT);
template <class M>
cudaError_t launch_nuts_piped(int NIT, const KParams<M>& P, size_t smem_per_chain, size_t block_data,
                              int cpb, int grid, cudaStream_t stream);
template <class M>
cudaError_t launch_nuts_lr(int W, const KParams<M>& P, size_t smem_per_chain, size_t block_data,
                           int cpb, int grid, int block, cudaStream_t stream);
template <class M>
size_t model_block_data_bytes(const typename M::Data& md);
template <class M>
cudaError_t launch_component(int W, const KParams<M>& P, int mode, const double* scal, double* out,
                             size_t smem, unsigned n);
template <class M>
size_t smem_fixed(int W, const typename M::Data& md, int Dp);

// NB200_MODEL_CUSTOM (kernels_custom.cu): register a CUDA source (identical text -> same slot),
// compile it with NVRTC for one geometry (no GPU needed), text of the last failure on this thread
int custom_register(const char* source);
int custom_compile(int program, int W, int NIT);
const char* custom_last_log();

}  // namespace nb200
