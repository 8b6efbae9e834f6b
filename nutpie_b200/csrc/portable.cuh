// portable.cuh — the few macros that let the sampler core compile both as
// sm_100a device code (the product) and, for tests/emul only, as plain host C++
// with a one-thread "group" so the control flow can be checked without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#ifndef __CUDACC__
#include <cmath>
#endif

#ifdef __CUDACC__
#define NB_HD __host__ __device__ __forceinline__
#define NB_D __device__ __forceinline__
#else
#define NB_HD inline
#define NB_D inline
#endif

#ifndef __CUDACC__
struct alignas(16) double2 {
    double x, y;
};
#endif

namespace nb200 {

NB_HD void nb_sincos_2pi(double u, double& s, double& c) {
#ifdef __CUDA_ARCH__
    sincospi(2.0 * u, &s, &c);
#else
    double t = 6.283185307179586476925286766559 * u;
    s = sin(t);
    c = cos(t);
#endif
}

template <class T>
NB_HD T nb_ldg(const T* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

NB_HD int nb_ffsll(unsigned long long x) {
#ifdef __CUDA_ARCH__
    return __ffsll((long long)x);
#else
    return __builtin_ffsll((long long)x);
#endif
}

NB_HD void nb_threadfence() {
#ifdef __CUDA_ARCH__
    __threadfence();
#endif
}

NB_HD void nb_threadfence_system() {
#ifdef __CUDA_ARCH__
    __threadfence_system();
#endif
}

NB_HD bool nb_isfinite(double x) {
#ifdef __CUDA_ARCH__
    return isfinite(x);
#else
    return std::isfinite(x);
#endif
}

// ---- bulk-copy staging (TMA 1-D bulk copies completing on an mbarrier) ----
// Used by the streaming leapfrog: one elected thread asks the copy engine for the next
// chunk of every vector, so the bytes in flight no longer depend on registers per thread.
#ifdef __CUDACC__
NB_D uint32_t nb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
NB_D void nb_mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nb_smem_u32(bar)), "r"(count) : "memory");
}
NB_D void nb_mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// (shared-memory operands are 32-bit shared-window addresses, computed once per pass)
NB_D void nb_mbar_expect_tx(uint32_t bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
NB_D void nb_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
NB_D void nb_mbar_wait(uint32_t bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned
NB_D void nb_bulk_g2s(uint32_t dst_smem, const void* src_gmem, unsigned bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dst_smem),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}
// the same copy with an L2 eviction-priority hint (createpolicy handle)
NB_D void nb_bulk_g2s_hint(uint32_t dst_smem, const void* src_gmem, unsigned bytes, uint32_t bar,
                           unsigned long long policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            dst_smem),
        "l"(src_gmem), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
// L2 eviction-priority policies: streamed-once data leaves first, re-read tables stay
NB_D unsigned long long nb_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
NB_D unsigned long long nb_policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// order this thread's earlier generic-proxy writes before later async-proxy (bulk copy) reads
NB_D void nb_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
#endif

}  // namespace nb200
