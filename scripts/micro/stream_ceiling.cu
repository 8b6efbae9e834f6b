// stream_ceiling.cu — what HBM bandwidth does the ACCESS PATTERN of the streaming leapfrog
// reach when nothing but the copies is left?  One CTA per "chain" (512 x 256 threads, 4 per
// SM), each pass reads q, p, s of one pool slot + the chain's mass matrix (4 x 80 KB) through
// double-buffered 1-D bulk copies and writes q', p', s' (3 x 80 KB) to the next slot — exactly
// the traffic of nuts_core.cuh's elementwise leapfrog, without tree logic or reductions.
// Variants: number of stages, chunk size, plain st.global vs bulk (TMA) stores from shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_ceiling stream_ceiling.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* b, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect(void* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(void* b, unsigned parity) {
    unsigned ok; const unsigned a = s32(b);
    do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(a), "r"(parity) : "memory"); } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// STAGES x 4 vectors x CHUNK double2.  TMA_ST: results go back into the stage (in place of the
// inputs) and leave through bulk stores.  EXTRA: extra read-only partner vectors per pass (0/2).
template <int STAGES, int CHUNK, bool TMA_ST, int T>
__global__ void __launch_bounds__(T, 1024 / T)
stream_kernel(double* pool, const double* var, int D2, int NS, int passes, double eps, double* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int VEC = CHUNK * 16;
    constexpr int STAGE = 4 * VEC;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + STAGES * STAGE);
    const int tid = threadIdx.x;
    const size_t Dp = (size_t)D2 * 2;
    double* mypool = pool + (size_t)blockIdx.x * NS * 3 * Dp;
    const double2* vr = reinterpret_cast<const double2*>(var + (size_t)blockIdx.x * Dp);
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned phase = 0;
    const int nchunk = (D2 + CHUNK - 1) / CHUNK;
    double acc = 0.0;
    for (int it = 0; it < passes; ++it) {
        const int src = it % NS, dst = (it + 1) % NS;
        const bool rev = it & 1;
        const double2* q = reinterpret_cast<const double2*>(mypool + (size_t)src * 3 * Dp);
        const double2* p = q + D2;
        const double2* s = p + D2;
        double2* qd = reinterpret_cast<double2*>(mypool + (size_t)dst * 3 * Dp);
        double2* pd = qd + D2;
        double2* sd = pd + D2;
        auto issue = [&](int c) {
            const int st = c % STAGES;
            const int first = (rev ? nchunk - 1 - c : c) * CHUNK;
            const int n = (D2 - first) < CHUNK ? (D2 - first) : CHUNK;
            const unsigned bytes = n * 16u;
            unsigned char* b = smem + st * STAGE;
            if (TMA_ST) bulk_wait_read<0>();  // the stage's previous results have left shared memory
            mbar_expect(bars + st, 4 * bytes);
            bulk_g2s(b, q + first, bytes, bars + st);
            bulk_g2s(b + VEC, p + first, bytes, bars + st);
            bulk_g2s(b + 2 * VEC, vr + first, bytes, bars + st);
            bulk_g2s(b + 3 * VEC, s + first, bytes, bars + st);
        };
        fence_async();
        __syncthreads();
        if (tid == 0)
            for (int c = 0; c < STAGES && c < nchunk; ++c) issue(c);
        for (int c = 0; c < nchunk; ++c) {
            const int st = c % STAGES;
            mbar_wait(bars + st, (phase >> st) & 1u);
            phase ^= 1u << st;
            const int first = (rev ? nchunk - 1 - c : c) * CHUNK;
            double2* b = reinterpret_cast<double2*>(smem + st * STAGE);
#pragma unroll
            for (int u = 0; u < CHUNK / T; ++u) {
                const int j = tid + u * T;
                if (first + j < D2) {
                    const double2 q0 = b[j], p0 = b[CHUNK + j], v0 = b[2 * CHUNK + j], s0 = b[3 * CHUNK + j];
                    double2 qn, pn, sn;
                    pn.x = p0.x - eps * q0.x; pn.y = p0.y - eps * q0.y;
                    qn.x = q0.x + eps * v0.x * pn.x; qn.y = q0.y + eps * v0.y * pn.y;
                    sn.x = s0.x + pn.x; sn.y = s0.y + pn.y;
                    acc += pn.x * v0.x * pn.x + pn.y * v0.y * pn.y;
                    if (TMA_ST) {
                        b[j] = qn; b[CHUNK + j] = pn; b[3 * CHUNK + j] = sn;
                    } else {
                        qd[first + j] = qn; pd[first + j] = pn; sd[first + j] = sn;
                    }
                }
            }
            if (TMA_ST) {
                fence_async();
                __syncthreads();
                if (tid == 0) {
                    const int n = (D2 - first) < CHUNK ? (D2 - first) : CHUNK;
                    const unsigned bytes = n * 16u;
                    unsigned char* bb = smem + st * STAGE;
                    bulk_s2g(qd + first, bb, bytes);
                    bulk_s2g(pd + first, bb + VEC, bytes);
                    bulk_s2g(sd + first, bb + 3 * VEC, bytes);
                    bulk_commit();
                    if (c + STAGES < nchunk) issue(c + STAGES);
                }
            } else if (c + STAGES < nchunk) {
                __syncthreads();
                if (tid == 0) issue(c + STAGES);
            }
        }
        if (TMA_ST && tid == 0) bulk_wait_read<0>();
        if (TMA_ST) { if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
        __syncthreads();
    }
    if (acc == 1.2345e-300) sink[0] = acc;
}

// plain LDG/STG version for reference (what the kernel did before staging)
__global__ void __launch_bounds__(256, 4)
ldg_kernel(double* pool, const double* var, int D2, int NS, int passes, double eps, double* sink) {
    const int tid = threadIdx.x;
    const size_t Dp = (size_t)D2 * 2;
    double* mypool = pool + (size_t)blockIdx.x * NS * 3 * Dp;
    const double2* vr = reinterpret_cast<const double2*>(var + (size_t)blockIdx.x * Dp);
    double acc = 0.0;
    for (int it = 0; it < passes; ++it) {
        const int src = it % NS, dst = (it + 1) % NS;
        const double2* q = reinterpret_cast<const double2*>(mypool + (size_t)src * 3 * Dp);
        const double2* p = q + D2; const double2* s = p + D2;
        double2* qd = reinterpret_cast<double2*>(mypool + (size_t)dst * 3 * Dp);
        double2* pd = qd + D2; double2* sd = pd + D2;
        for (int k = tid; k < D2; k += 256) {
            const double2 q0 = q[k], p0 = p[k], v0 = vr[k], s0 = s[k];
            double2 qn, pn, sn;
            pn.x = p0.x - eps * q0.x; pn.y = p0.y - eps * q0.y;
            qn.x = q0.x + eps * v0.x * pn.x; qn.y = q0.y + eps * v0.y * pn.y;
            sn.x = s0.x + pn.x; sn.y = s0.y + pn.y;
            acc += pn.x * v0.x * pn.x + pn.y * v0.y * pn.y;
            qd[k] = qn; pd[k] = pn; sd[k] = sn;
        }
        __syncthreads();
    }
    if (acc == 1.2345e-300) sink[0] = acc;
}

template <int STAGES, int CHUNK, bool TMA_ST, int T>
static void run(const char* name, double* pool, double* var, double* sink, int chains, int D, int NS, int passes) {
    const int D2 = D / 2;
    const size_t smem = (size_t)STAGES * 4 * CHUNK * 16 + 8 * STAGES;
    auto k = stream_kernel<STAGES, CHUNK, TMA_ST, T>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, T, smem));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a));
        k<<<chains, T, smem>>>(pool, var, D2, NS, passes, 1e-3, sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaGetLastError());
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double bytes = 56.0 * D * (double)passes * chains;
    printf("%-44s occ %d/SM  %8.2f ms  %7.1f GB/s moved  (%5.1f%% of 6555)  = %.3g passes/s\n", name, occ, best,
           bytes / best / 1e6, bytes / best / 1e6 / 65.55, (double)passes * chains / best * 1e3);
    fflush(stdout);
}

int main(int argc, char** argv) {
    const int chains = argc > 1 ? atoi(argv[1]) : 512;
    const int D = argc > 2 ? atoi(argv[2]) : 10000;
    const int passes = argc > 3 ? atoi(argv[3]) : 300;
    const int NS = 36;
    const size_t Dp = D;
    double *pool, *var, *sink;
    CK(cudaMalloc(&pool, sizeof(double) * chains * NS * 3 * Dp));
    CK(cudaMalloc(&var, sizeof(double) * chains * Dp));
    CK(cudaMalloc(&sink, 64));
    CK(cudaMemset(pool, 0, sizeof(double) * chains * NS * 3 * Dp));
    CK(cudaMemset(var, 0, sizeof(double) * chains * Dp));
    printf("chains %d  D %d  passes %d  pool %.2f GB\n", chains, D, passes, chains * NS * 3.0 * Dp * 8 / 1e9);
    {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a);
            ldg_kernel<<<chains, 256>>>(pool, var, D / 2, NS, passes, 1e-3, sink);
            cudaEventRecord(b); CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        const double bytes = 56.0 * D * (double)passes * chains;
        printf("%-44s            %8.2f ms  %7.1f GB/s moved  (%5.1f%% of 6555)\n", "ldg/stg 256 thr", best, bytes / best / 1e6, bytes / best / 1e6 / 65.55);
    }
    run<2, 256, false, 256>("tma-load 2 stages x256, st.global", pool, var, sink, chains, D, NS, passes);
    run<3, 256, false, 256>("tma-load 3 stages x256, st.global", pool, var, sink, chains, D, NS, passes);
    run<2, 256, true, 256>("tma-load 2 stages x256, tma-store", pool, var, sink, chains, D, NS, passes);
    run<3, 256, true, 256>("tma-load 3 stages x256, tma-store", pool, var, sink, chains, D, NS, passes);
    run<2, 512, false, 256>("tma-load 2 stages x512, st.global (3/SM)", pool, var, sink, chains, D, NS, passes);
    run<2, 512, true, 256>("tma-load 2 stages x512, tma-store (3/SM)", pool, var, sink, chains, D, NS, passes);
    run<4, 128, false, 128>("128 thr: 4 stages x128, st.global", pool, var, sink, chains, D, NS, passes);
    run<4, 128, true, 128>("128 thr: 4 stages x128, tma-store", pool, var, sink, chains, D, NS, passes);
    run<6, 128, true, 128>("128 thr: 6 stages x128, tma-store", pool, var, sink, chains, D, NS, passes);
    run<3, 512, true, 512>("512 thr: 3 stages x512, tma-store (2/SM)", pool, var, sink, chains, D, NS, passes);
    return 0;
}
