// d2h_ceiling.cu — what device-to-host bandwidth does this BOX sustain when n GPUs copy into
// pinned host memory at the same time?  The end-to-end arm of bench.py lands 5.9 GB of trace per
// GPU and step in host memory; at 8 GPUs its scaling is bounded by this number, not by the
// sampler (e2e efficiency 0.43 at N = 8).  One host thread + stream per device,
// 256 MB chunks, cudaMemcpyAsync from device memory to cudaHostAlloc'ed buffers, 2 s per point.
// Build: nvcc -O2 -o d2h_ceiling d2h_ceiling.cu -lpthread      Usage: ./d2h_ceiling [max_gpus]
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/mman.h>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

int main(int argc, char** argv) {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (argc > 1 && atoi(argv[1]) < ndev) ndev = atoi(argv[1]);
    const size_t chunk = 256ull << 20, ring = 4;  // 1 GB of pinned memory per device
    std::vector<void*> dsrc(ndev), hdst(ndev);
    std::vector<cudaStream_t> st(ndev);
    for (int d = 0; d < ndev; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaMalloc(&dsrc[d], chunk));
        CK(cudaMemset(dsrc[d], 1, chunk));
        if (argc > 2 && !strcmp(argv[2], "huge")) {
            // destination in transparent huge pages, pinned by registration: does the page size
            // of the host buffer change what the box's DMA path sustains?
            if (posix_memalign(&hdst[d], 2u << 20, chunk * ring)) { printf("posix_memalign failed\n"); return 1; }
            madvise(hdst[d], chunk * ring, MADV_HUGEPAGE);
            memset(hdst[d], 0, chunk * ring);
            CK(cudaHostRegister(hdst[d], chunk * ring, cudaHostRegisterPortable));
        } else {
            CK(cudaHostAlloc(&hdst[d], chunk * ring, cudaHostAllocDefault));
        }
        CK(cudaStreamCreateWithFlags(&st[d], cudaStreamNonBlocking));
    }
    printf("{\"buffers\": \"%s\", \"how\": \"n devices copy 256 MB chunks device->pinned host concurrently for 2 s each\", \"points\": [", argc > 2 ? argv[2] : "cudaHostAlloc");
    for (int n = 1; n <= ndev; n = n < 4 ? n * 2 : n + 2) {
        std::atomic<bool> go{false}, stop{false};
        std::vector<double> bytes(n, 0.0);
        std::vector<std::thread> th;
        for (int d = 0; d < n; ++d)
            th.emplace_back([&, d] {
                cudaSetDevice(d);
                while (!go.load()) std::this_thread::yield();
                size_t i = 0;
                while (!stop.load()) {
                    cudaMemcpyAsync((char*)hdst[d] + (i % ring) * chunk, dsrc[d], chunk, cudaMemcpyDeviceToHost, st[d]);
                    cudaStreamSynchronize(st[d]);
                    bytes[d] += (double)chunk;
                    ++i;
                }
            });
        auto t0 = std::chrono::steady_clock::now();
        go.store(true);
        std::this_thread::sleep_for(std::chrono::seconds(2));
        stop.store(true);
        for (auto& t : th) t.join();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        double tot = 0.0, mn = 1e30;
        for (double b : bytes) { tot += b; if (b < mn) mn = b; }
        printf("%s{\"gpus\": %d, \"aggregate_gbs\": %.1f, \"slowest_gpu_gbs\": %.1f}", n > 1 ? ", " : "", n,
               tot / dt / 1e9, mn / dt / 1e9);
        fflush(stdout);
    }
    printf("], \"host_copy\": [");
    // what the host cores themselves sustain: T threads, each streaming 64 MB blocks src -> dst
    // (the cost of expanding / regrouping a trace on the host instead of on the device)
    {
        const size_t blk = 64ull << 20;
        const int maxT = (int)std::thread::hardware_concurrency();
        bool first = true;
        for (int T = 1; T <= maxT; T *= 2) {
            std::vector<char*> src(T), dst(T);
            for (int t = 0; t < T; ++t) {
                src[t] = (char*)malloc(blk);
                dst[t] = (char*)malloc(blk);
                memset(src[t], 1, blk);
                memset(dst[t], 2, blk);
            }
            std::atomic<bool> go{false}, stop{false};
            std::vector<double> bytes(T, 0.0);
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t)
                th.emplace_back([&, t] {
                    while (!go.load()) std::this_thread::yield();
                    while (!stop.load()) {
                        memcpy(dst[t], src[t], blk);
                        bytes[t] += (double)blk;
                    }
                });
            auto t0 = std::chrono::steady_clock::now();
            go.store(true);
            std::this_thread::sleep_for(std::chrono::milliseconds(700));
            stop.store(true);
            for (auto& t : th) t.join();
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            double tot = 0.0;
            for (double b : bytes) tot += b;
            printf("%s{\"threads\": %d, \"copied_gbs\": %.1f}", first ? "" : ", ", T, tot / dt / 1e9);
            first = false;
            fflush(stdout);
            for (int t = 0; t < T; ++t) { free(src[t]); free(dst[t]); }
        }
    }
    printf("]}\n");
    return 0;
}
