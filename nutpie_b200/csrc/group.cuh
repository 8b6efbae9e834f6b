// group.cuh — the set of threads that cooperates on ONE chain.
//
// nuts-rs runs one chain per rayon task and vectorises inside the chain with
// pulp SIMD (its `CpuMath`, instantiated at src/pymc.rs:496-503).  On B200 a
// chain is owned by W warps (W = 1: a warp per chain, several chains per CTA;
// W > 1: a CTA per chain).  All scalar control flow of the chain is executed
// redundantly by every thread of the group; reductions return bit-identical
// values on every thread (xor butterfly + fixed-order cross-warp sum), so the
// control flow stays uniform without broadcasts.
#pragma once
#include "portable.cuh"

namespace nb200 {

#ifdef __CUDACC__
template <int W>
struct GroupCuda {
    static constexpr int kThreads = 32 * W;
    int tid;       // thread index inside the chain's group
    double* red;   // shared scratch, W * kMaxRed doubles (only W > 1)
    static constexpr int kMaxRed = 12;

    NB_D int size() const { return kThreads; }
    NB_D void sync() const {
        if (W == 1) __syncwarp();
        else __syncthreads();
    }
    // sum N values over the group; result identical on all threads
    template <int N>
    NB_D void reduce(double (&v)[N]) const {
#pragma unroll
        for (int i = 0; i < N; ++i) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
        }
        if (W > 1) {
            const int warp = tid >> 5;
            __syncthreads();  // previous users of `red` are done
            if ((tid & 31) == 0) {
#pragma unroll
                for (int i = 0; i < N; ++i) red[warp * kMaxRed + i] = v[i];
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < W; ++w) s += red[w * kMaxRed + i];
                v[i] = s;
            }
        }
    }
};
#endif

// one host thread plays the whole group (tests/emul only)
struct GroupSerial {
    static constexpr int kThreads = 1;
    int tid;
    NB_HD int size() const { return 1; }
    NB_HD void sync() const {}
    template <int N>
    NB_HD void reduce(double (&)[N]) const {}
};

}  // namespace nb200
