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

NB_HD bool nb_isfinite(double x) {
#ifdef __CUDA_ARCH__
    return isfinite(x);
#else
    return std::isfinite(x);
#endif
}

}  // namespace nb200
