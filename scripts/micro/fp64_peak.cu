// fp64_peak.cu — measured FP64 FMA throughput of this GPU, the denominator of the radon
// kernel's roofline (bench.py `roofline.bound = "fp64_issue"`): MEASURED_PEAKS.json has no
// FP64 figure, and the nominal 40 TFLOP/s is a marketing number, so it is measured here the
// same way the driver measures its bf16 peak (best of several timed launches, CUDA events).
// Every thread runs ILP independent DFMA chains; the grid fills every SM with 2048 threads.
// Also reports the rate one warp per scheduler reaches with ONE dependent chain (the
// latency-bound regime the radon kernel lives in: 1.7 warps per scheduler).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(1024) fma_kernel(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = (double)threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;  // never true: keeps the chains alive
}

template <int ILP>
static double run(int grid, int block, int iters) {
    double* out;
    CK(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    fma_kernel<ILP><<<grid, block>>>(out, iters, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        fma_kernel<ILP><<<grid, block>>>(out, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * ILP * (double)iters * grid * (double)block;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    CK(cudaFree(out));
    return best;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz\": %d,\n", p.name, sms, p.clockRate / 1000);
    // saturated: 2 CTAs x 1024 threads per SM, 8 independent chains per thread
    const double sat8 = run<8>(sms * 2, 1024, 20000);
    const double sat4 = run<4>(sms * 2, 1024, 20000);
    // latency-bound: 7 warps per SM (the radon geometry), one / four chains per thread
    const double lat1 = run<1>(sms, 224, 200000);
    const double lat4 = run<4>(sms, 224, 100000);
    printf(" \"fp64_fma_tflops_saturated\": %.2f, \"fp64_fma_tflops_saturated_ilp4\": %.2f,\n", sat8, sat4);
    printf(" \"fp64_fma_tflops_7warps_per_sm_ilp1\": %.3f, \"fp64_fma_tflops_7warps_per_sm_ilp4\": %.3f,\n", lat1, lat4);
    // dependent-issue latency of DFMA in cycles: 7 warps/SM never contend for the pipe at ILP 1
    const double fma_per_s_per_warp = lat1 * 1e12 / 2.0 / 32.0 / (sms * 7.0);
    printf(" \"dfma_dependent_latency_cycles\": %.2f,\n", (p.clockRate * 1e3) / fma_per_s_per_warp);
    printf(" \"how\": \"independent DFMA chains, best of 5 launches, CUDA events; flops = 2 per FMA\"}\n");
    return 0;
}
