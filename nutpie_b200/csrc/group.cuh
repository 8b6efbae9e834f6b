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
    // Warp sum of N values with recursive halving: at every level a lane keeps half of its
    // values and hands the other half to its xor-partner, so the shuffle count is
    // ~N + log2(32) instead of N * log2(32); the final value i lives in the lanes whose
    // low bits spell i and is broadcast from the lowest such lane, so every lane ends up with
    // bit-identical sums (fixed association order).
    template <int N>
    NB_D static void warp_reduce(double (&v)[N]) {
        if constexpr (N < 4) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
            }
        } else {
            constexpr int P = N <= 4 ? 4 : (N <= 8 ? 8 : 16);  // padded to a power of two
            constexpr int LV = P == 4 ? 2 : (P == 8 ? 3 : 4);  // halving levels
            const int lane = threadIdx.x & 31;
            double w[P];
#pragma unroll
            for (int i = 0; i < P; ++i) w[i] = i < N ? v[i] : 0.0;
            // level l pairs lanes differing in bit l; the lane with bit l set keeps the upper half
            int cnt = P;
#pragma unroll
            for (int l = 0; l < LV; ++l) {
                const int half = cnt >> 1;
                const bool up = (lane >> l) & 1;
#pragma unroll
                for (int i = 0; i < P / 2; ++i) {
                    if (i < half) {
                        const double keep = up ? w[half + i] : w[i];
                        const double give = up ? w[i] : w[half + i];
                        w[i] = keep + __shfl_xor_sync(0xffffffffu, give, 1 << l);
                    }
                }
                cnt = half;
            }
            // w[0] now holds the partial sum of value id(lane) = bit-reversed low LV bits order
            double r = w[0];
#pragma unroll
            for (int off = 1 << LV; off < 32; off <<= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
            // value index held by a lane: bit l of lane selects the upper half at level l
#pragma unroll
            for (int i = 0; i < N; ++i) {
                int src = 0, lo = 0, span = P;
#pragma unroll
                for (int l = 0; l < LV; ++l) {
                    span >>= 1;
                    if (i >= lo + span) {
                        src |= 1 << l;
                        lo += span;
                    }
                }
                v[i] = __shfl_sync(0xffffffffu, r, src);
            }
        }
    }

    // exclusive prefix sum over the group in thread order: thread t gets x_0 + ... + x_{t-1}
    // (Hillis-Steele inside a warp, then the earlier warps' totals added in warp order)
    NB_D double exclusive_scan(double x) const {
        const int lane = tid & 31;
        double v = x;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, v, off);
            if (lane >= off) v += y;
        }
        double excl = __shfl_up_sync(0xffffffffu, v, 1);
        if (lane == 0) excl = 0.0;
        if (W > 1) {
            const int warp = tid >> 5;
            __syncthreads();  // previous users of `red` are done
            if (lane == 31) red[warp] = v;
            __syncthreads();
            double base = 0.0;
            for (int w = 0; w < warp; ++w) base += red[w];
            excl = base + excl;
        }
        return excl;
    }

    // sum N values over the group; result identical on all threads
    template <int N>
    NB_D void reduce(double (&v)[N]) const {
        warp_reduce(v);
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

#ifdef __CUDACC__
// L lanes of a warp own one chain (L = 4, 8, 16): a warp hosts 32 / L chains.  For densities with
// a handful of dimensions (BASELINE configs 1 and 5: D = 1, 9) a full warp per chain leaves most
// lanes idle on every instruction; here the chains of a warp run the same code side by side and
// only part ways where their trees differ (independent thread scheduling serialises the
// divergent stretches and reconverges them).  Reductions are xor butterflies inside the aligned
// L-lane group, so every lane of a chain holds bit-identical sums.
template <int L>
struct GroupSub {
    static_assert(L == 4 || L == 8 || L == 16, "sub-warp groups are 4, 8 or 16 lanes");
    static constexpr int kThreads = L;
    int tid;        // lane inside the chain's group
    unsigned mask;  // the group's lanes inside the warp
    double* red;    // unused (interface parity with GroupCuda)
    static constexpr int kMaxRed = 12;

    NB_D int size() const { return L; }
    NB_D void sync() const { __syncwarp(mask); }
    template <int N>
    NB_D void reduce(double (&v)[N]) const {
#pragma unroll
        for (int i = 0; i < N; ++i) {
#pragma unroll
            for (int off = L / 2; off > 0; off >>= 1) v[i] += __shfl_xor_sync(mask, v[i], off);
        }
    }
    NB_D double exclusive_scan(double x) const {
        double v = x;
#pragma unroll
        for (int off = 1; off < L; off <<= 1) {
            const double y = __shfl_up_sync(mask, v, off, L);
            if (tid >= off) v += y;
        }
        double excl = __shfl_up_sync(mask, v, 1, L);
        if (tid == 0) excl = 0.0;
        return excl;
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
    NB_HD double exclusive_scan(double) const { return 0.0; }
};

}  // namespace nb200
