// nb200_api.cu — sm_100a kernels + the C-ABI of include/nutpie_b200.h.
//
// One persistent kernel launch advances every chain of a sampler through all
// of its tuning and sampling draws (nuts_kernel); chains never synchronise with
// each other.  The host side of this file is the replacement for what
// nuts_rs::Sampler does around its rayon pool (src/wrapper.rs:977-1456):
// start, poll progress, pause/resume/abort, hand out the trace.
#include <cuda_runtime.h>

#include <sched.h>
#include <sys/mman.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "launch.hpp"
#include "nuts_core.cuh"
#include "radon_layout.hpp"

using namespace nb200;

// the Python/Rust bindings mirror these layouts field by field
static_assert(sizeof(nb200_settings) == 240, "nb200_settings ABI layout changed");
static_assert(sizeof(nb200_model_desc) == 144, "nb200_model_desc ABI layout changed");
static_assert(sizeof(nb200_progress) == 56, "nb200_progress ABI layout changed");

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                        \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess)                                                          \
            return fail(NB200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define CUP(call)                                                                       \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            fail(NB200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
            return nullptr;                                                             \
        }                                                                               \
    } while (0)

static std::atomic<int> g_threads_per_chain{0};
static std::atomic<int> g_chains_per_block{0};
static std::atomic<int> g_smem_slots{-1};
static std::atomic<int> g_force_nit{-1};
static std::atomic<int> g_pipeline{0};     // two-warp producer / consumer kernels: opt-in (nb200_set_pipeline)
static std::atomic<int> g_stage_loads{3};  // staging + alternating sweep (nb200_set_stage_loads)
static inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
// + working mass matrix + hot tier of the pool
static size_t chain_smem_total(size_t fixed, int Dp, int var_in_smem, int smem_slots) {
    return fixed + (var_in_smem ? align16(sizeof(double) * Dp) : 0) +
           align16(sizeof(double) * 4 * (size_t)Dp * smem_slots);
}

// Device buffers of a sampler come from the stream-ordered memory pool with an unbounded
// release threshold: destroying a sampler returns its ~3.4 GB to the pool instead of
// unmapping it (cudaFree of that much memory was measured at 100-180 ms), and the next
// sampler's allocations are served from the pool.
static void pool_keep_memory(int device) {
    static std::mutex mu;
    static std::vector<int> done;
    std::lock_guard<std::mutex> lk(mu);
    for (int d : done)
        if (d == device) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done.push_back(device);
}


// ------------------------------------------------------------------ host plug-in service
// NB200_MODEL_HOST: the reference's RawLogpFunc (src/pymc.rs:23-29) is called by a pool of host
// threads — the role nuts-rs gives its rayon workers (src/wrapper.rs:977) — on behalf of the
// chains running in the persistent kernel.  Chain c posts q to qbox[c], rings req[c] (models.cuh,
// HostModel::logp_grad); the thread that owns c calls the pointer, writes gbox[c] / lpbox[c] /
// rcbox[c] and publishes resp[c] = req[c].  All six arrays and the stop flag live in MAPPED
// pinned memory, so neither side issues a copy.
extern "C" {
static bool is_pinned_host(const void* p);
}
static std::atomic<int> g_persist_users[16];  // samplers holding a persisting L2 window, per device
static int usable_cores() {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        const int n = CPU_COUNT(&set);
        if (n > 0) return n;
    }
    const unsigned h = std::thread::hardware_concurrency();
    return h ? (int)h : 1;
}

struct HostService {
    nb200_logp_fn fn = nullptr;
    const void* user_data = nullptr;
    size_t D = 0, Dp = 0;
    uint64_t n_chains = 0;
    void* base = nullptr;  // one mapped pinned allocation
    double *qbox = nullptr, *gbox = nullptr, *lpbox = nullptr;
    int* rcbox = nullptr;
    unsigned *req = nullptr, *resp = nullptr;
    int* stop = nullptr;
    // the same arrays as the device sees them
    double *d_qbox = nullptr, *d_gbox = nullptr, *d_lpbox = nullptr;
    int* d_rcbox = nullptr;
    unsigned *d_req = nullptr, *d_resp = nullptr;
    int* d_stop = nullptr;
    std::vector<std::thread> threads;
    std::atomic<bool> quit{false};
    std::atomic<int> fatal_rc{0};
    std::atomic<long long> fatal_chain{-1};
    std::atomic<unsigned long long> calls{0};

    int init(const nb200_model_desc& m, uint64_t chains, int dp) {
        fn = m.host_logp;
        user_data = m.host_user_data;
        D = (size_t)m.dim;
        Dp = (size_t)dp;
        n_chains = chains;
        const size_t vec = align16(sizeof(double) * Dp * n_chains);
        const size_t sc = align16(sizeof(double) * n_chains);
        const size_t wd = align16(sizeof(unsigned) * n_chains);
        const size_t total = 2 * vec + sc + 3 * wd + 64;
        if (cudaHostAlloc(&base, total, cudaHostAllocMapped) != cudaSuccess)
            return fail(NB200_ECUDA, "cudaHostAlloc (mapped mailbox of the host plug-in) failed");
        std::memset(base, 0, total);
        char* b = static_cast<char*>(base);
        qbox = reinterpret_cast<double*>(b);
        gbox = reinterpret_cast<double*>(b + vec);
        lpbox = reinterpret_cast<double*>(b + 2 * vec);
        rcbox = reinterpret_cast<int*>(b + 2 * vec + sc);
        req = reinterpret_cast<unsigned*>(b + 2 * vec + sc + wd);
        resp = reinterpret_cast<unsigned*>(b + 2 * vec + sc + 2 * wd);
        stop = reinterpret_cast<int*>(b + 2 * vec + sc + 3 * wd);
        void* dbase = nullptr;
        if (cudaHostGetDevicePointer(&dbase, base, 0) != cudaSuccess)
            return fail(NB200_ECUDA, "cudaHostGetDevicePointer failed");
        char* d = static_cast<char*>(dbase);
        d_qbox = reinterpret_cast<double*>(d);
        d_gbox = reinterpret_cast<double*>(d + vec);
        d_lpbox = reinterpret_cast<double*>(d + 2 * vec);
        d_rcbox = reinterpret_cast<int*>(d + 2 * vec + sc);
        d_req = reinterpret_cast<unsigned*>(d + 2 * vec + sc + wd);
        d_resp = reinterpret_cast<unsigned*>(d + 2 * vec + sc + 2 * wd);
        d_stop = reinterpret_cast<int*>(d + 2 * vec + sc + 3 * wd);
        int T = m.host_threads > 0 ? m.host_threads : usable_cores();
        if ((uint64_t)T > n_chains) T = (int)n_chains;
        if (T < 1) T = 1;
        for (int t = 0; t < T; ++t) threads.emplace_back([this, t, T] { worker(t, T); });
        return 0;
    }

    void worker(int t, int T) {
        // a contiguous block of chains per thread: its doorbells share cache lines
        const uint64_t lo = n_chains * (uint64_t)t / T, hi = n_chains * (uint64_t)(t + 1) / T;
        std::vector<unsigned> last(hi - lo, 0u);
        std::vector<double> grad(D);
        unsigned idle = 0;
        while (!quit.load(std::memory_order_relaxed)) {
            bool any = false;
            for (uint64_t c = lo; c < hi; ++c) {
                const unsigned r = __atomic_load_n(req + c, __ATOMIC_ACQUIRE);
                if (r == last[c - lo]) continue;
                any = true;
                double lp = 0.0;
                int rc = fn(D, qbox + c * Dp, grad.data(), &lp, user_data);
                // the reference's cfunc wrapper reports non-finite values itself
                // (compile_pymc.py:996-999); do the same for pointers that do not
                if (rc == 0) {
                    if (!std::isfinite(lp)) rc = 4;
                    else
                        for (size_t i = 0; i < D; ++i)
                            if (!std::isfinite(grad[i])) { rc = 3; break; }
                }
                std::memcpy(gbox + c * Dp, grad.data(), sizeof(double) * D);
                lpbox[c] = lp;
                rcbox[c] = rc;
                if (rc < 0) {  // fatal (src/pymc.rs:178): stop the sampler, keep the partial trace
                    int expect = 0;
                    if (fatal_rc.compare_exchange_strong(expect, rc)) fatal_chain.store((long long)c);
                    __atomic_store_n(stop, 1, __ATOMIC_RELEASE);
                }
                calls.fetch_add(1, std::memory_order_relaxed);
                last[c - lo] = r;
                __atomic_store_n(resp + c, r, __ATOMIC_RELEASE);
            }
            if (any) {
                idle = 0;
            } else if (++idle > 2000) {  // nothing posted for a while (paused / finished)
                std::this_thread::sleep_for(std::chrono::microseconds(200));
            } else if (idle > 50) {
                std::this_thread::yield();
            }
        }
    }

    void shutdown() {
        quit.store(true);
        for (auto& th : threads)
            if (th.joinable()) th.join();
        threads.clear();
        if (base) cudaFreeHost(base);
        base = nullptr;
    }
};

// ------------------------------------------------------------------ sampler
enum class RunState { Created, Running, Paused, Finished, Aborted, Error };

struct nb200_sampler {
    virtual ~nb200_sampler() {}
    int device = 0;
    nb200_settings st{};
    nb200_model_desc model{};
    uint64_t n_chains = 0, chain_id_offset = 0;
    int W = 1, NIT = 0, cpb = 1, grid = 0, block = 0;
    bool piped = false;  // two warps per chain: integrator + tree (nuts_kernel_piped)
    int sub = 0;         // lanes per chain of the sub-warp geometry (nuts_kernel_sub); 0 = none
    size_t smem_per_chain = 0, block_data = 0;
    cudaStream_t stream = nullptr, side = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    RunState state = RunState::Created;
    std::mutex mu;
    double kernel_ms = 0.0;
    uint64_t launches = 0;
    uint64_t draws_per_launch = 0;
    int* d_stop = nullptr;
    ChainScalars* d_sc = nullptr;
    ChainScalars* h_sc = nullptr;  // host copy of the progress records (pageable: pinning and
                                   // unpinning per sampler cost 10-400 ms in cudaHostAlloc/FreeHost)
    std::vector<ChainScalars> h_sc_store;
    double *d_pool = nullptr, *d_var = nullptr, *d_wf = nullptr;
    double *d_draws = nullptr, *d_stats = nullptr, *d_grads = nullptr, *d_mm = nullptr;
    double* d_div = nullptr;  // store_divergences: [n_rows][n_chains][4][grad_dim]
    // adaptation = low_rank (lowrank.cuh): per-chain metric, window, scratch; eigenvalue trace
    bool l2_persist = false;  // mass matrices held in a persisting L2 window (streaming regime)
    bool lr = false;
    int lr_cap = 0, lr_max_rank = 0;
    double *d_lr_stds = nullptr, *d_lr_vals = nullptr, *d_lr_vecs = nullptr, *d_lr_coef = nullptr;
    double *d_lr_win = nullptr, *d_lr_mat = nullptr, *d_lr_cols = nullptr, *d_eig = nullptr;
    std::unique_ptr<HostService> host;  // NB200_MODEL_HOST only
    std::string err;                    // message of the error that put the sampler in Error
    double *d_q0 = nullptr, *d_init_mean = nullptr, *d_tape = nullptr;
    // pinned host trace (lazy)
    double *h_draws = nullptr, *h_stats = nullptr, *h_grads = nullptr, *h_mm = nullptr;
    std::vector<uint64_t> rows_filled;
    uint64_t n_rows = 0, sdim = 0, n_total = 0;
    uint64_t grad_dim = 0;  // row width of the gradient / mass-matrix traces (never expanded)
    int Dp = 0, NS = 0, smem_slots = 0;
    std::vector<void*> model_allocs;
    bool model_allocs_pooled = false;
    virtual int launch() = 0;
    int sampler_error = 0;
    // streaming of finished trace rows to host buffers while the kernel runs
    double *tgt_draws = nullptr, *tgt_stats = nullptr;
    size_t tgt_draws_pitch = 0, tgt_stats_pitch = 0;  // row stride of the targets in doubles (0 = dense)
    uint64_t streamed_rows = 0;
    std::chrono::steady_clock::time_point last_stream{};
};

template <class M>
struct SamplerImpl : nb200_sampler {
    KParams<M> P;
    int launch() override {
        P.max_draws_per_launch = draws_per_launch;
        cudaError_t e =
            lr  ? launch_nuts_lr<M>(W, P, smem_per_chain, block_data, cpb, grid, block, stream)
            : sub ? launch_nuts_sub<M>(sub, NIT, P, smem_per_chain, block / 32, grid, stream)
            : piped ? launch_nuts_piped<M>(NIT, P, smem_per_chain, block_data, cpb, grid, stream)
                    : launch_nuts<M>(W, NIT, P, smem_per_chain, block_data, cpb, grid, block, stream);
        if (e != cudaSuccess) return fail(NB200_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
        return 0;
    }
};

template <class M>
static size_t smem_for(int W, const typename M::Data& md, int Dp) {
    return smem_fixed<M>(W, md, Dp);
}

static int pick_W(const nb200_model_desc& m, uint64_t n_chains) {
    int t = g_threads_per_chain.load();
    if (t == 4 || t == 8 || t == 16) return 1;  // sub-warp groups (create_impl picks the lanes)
    if (t > 0) return t / 32;
    if (m.dim >= 2048) {  // streaming regime (config 4): a CTA per chain; prefer CTAs small
        // enough that every chain is resident at once.  With bulk-copy staging the bytes in
        // flight do not depend on the thread count: 128 threads x 128 registers per chain beat
        // 256 x 64 (no spills in the streaming loop, half the redundant scalar work; measured
        // +13 %, profiles/r1_sweep_config4_variants.txt)
        if (n_chains > 148ull * 2) return m.kind == NB200_MODEL_NORMAL ? 4 : 8;
        if (n_chains > 148ull) return 16;
        return 32;
    }
    if (m.dim <= 256) return 1;    // a warp per chain with fully unrolled per-dimension loops
                                   // (measured best on radon: profiles/sweep_radon_r1.txt)
    const uint64_t work = m.kind == NB200_MODEL_RADON ? (uint64_t)m.n_obs : m.dim;
    // enough warps per chain that each thread still has >= 4 work items, and
    // enough warps in total to fill 148 SMs x 16 warps
    int w = 1;
    while (w < 32 && (uint64_t)(32 * w * 2) * 4 <= work && n_chains * (uint64_t)w < 148ull * 24) w *= 2;
    return w;
}

static int validate(const nb200_settings* st, const nb200_model_desc* m) {
    if (!st || !m) return fail(NB200_EINVAL, "null settings/model");
    if (st->maxdepth < 1 || 3 * ((int)st->maxdepth + 1) + 3 > kMaxSlots)
        return fail(NB200_EINVAL, "maxdepth must be in 1..19");
    if (m->dim < 1) return fail(NB200_EINVAL, "model dimension must be >= 1");
    if (m->dim > (1u << 22)) return fail(NB200_EINVAL, "model dimension must be <= 2^22");
    if (st->step_size_method < 0 || st->step_size_method > 2)
        return fail(NB200_EINVAL, "step_size_adapt_method must be dual_average (0), adam (1) or fixed (2)");
    if (st->step_size_method == 1 && !(st->adam_learning_rate > 0))
        return fail(NB200_EINVAL, "step_size_adam_learning_rate must be > 0");
    if (st->step_size_jitter < 0 || st->step_size_jitter >= 1)
        return fail(NB200_EINVAL, "step_size_jitter must be in [0, 1)");
    if (st->adaptation != 0 && st->adaptation != 1)
        return fail(NB200_EINVAL, "adaptation must be diag (0) or low_rank (1)");
    if (st->adaptation == 1) {
        // the refresh factorises two dim x dim matrices per chain (lowrank.cuh)
        if (m->dim > 1024)
            return fail(NB200_EINVAL, "adaptation='low_rank': the model dimension must be <= 1024 dimensions");
        if (!(st->mass_matrix_eigval_cutoff > 1.0))
            return fail(NB200_EINVAL, "mass_matrix_eigval_cutoff must be > 1");
        if (!(st->mass_matrix_gamma > 0.0)) return fail(NB200_EINVAL, "mass_matrix_gamma must be > 0");
        if (st->mass_matrix_max_rank < 1) return fail(NB200_EINVAL, "mass_matrix_max_rank must be >= 1");
    }
    if (m->kind == NB200_MODEL_RADON) {
        if (m->n_county < 1 || m->n_county > 32767)
            return fail(NB200_EINVAL, "radon: n_county must be in 1..32767");
        if (m->dim != (uint64_t)(2 * m->n_county + 5))
            return fail(NB200_EINVAL, "radon: dim must equal 2*n_county+5");
        if (!m->y || !m->county || !m->floor || m->n_obs < 1)
            return fail(NB200_EINVAL, "radon: missing data arrays");
        for (int i = 0; i < m->n_obs; ++i)
            if (m->county[i] < 0 || m->county[i] >= m->n_county)
                return fail(NB200_EINVAL, "radon: county index out of range");
    } else if (m->kind == NB200_MODEL_NORMAL) {
        if (!(m->sigma > 0)) return fail(NB200_EINVAL, "normal: sigma must be > 0");
    } else if (m->kind == NB200_MODEL_FUNNEL) {
        if (m->dim < 2) return fail(NB200_EINVAL, "funnel: dim must be >= 2");
    } else if (m->kind == NB200_MODEL_CUSTOM) {
        if (!m->cuda_source || !m->cuda_source[0])
            return fail(NB200_EINVAL, "custom: cuda_source is empty");
        if (m->n_user_data > 0 && !m->user_data)
            return fail(NB200_EINVAL, "custom: user_data is null but n_user_data > 0");
        if (m->n_user_data > (1ull << 31)) return fail(NB200_EINVAL, "custom: user_data too large");
        if (m->n_user_scratch > 16384) return fail(NB200_EINVAL, "custom: at most 16384 doubles of scratch");
    } else if (m->kind == NB200_MODEL_HOST) {
        if (!m->host_logp) return fail(NB200_EINVAL, "host: host_logp is null");
        if (m->host_threads < 0) return fail(NB200_EINVAL, "host: host_threads must be >= 0");
    } else {
        return fail(NB200_EINVAL, "unknown model kind");
    }
    return 0;
}

// small constant tables of a density: pool-backed (stream-ordered) when a stream is given
static thread_local cudaStream_t g_upload_stream = nullptr;
template <class T>
static int to_device(const std::vector<T>& v, T** out, std::vector<void*>& keep) {
    const size_t bytes = sizeof(T) * (v.size() ? v.size() : 1);
    if (g_upload_stream) {
        CU(cudaMallocAsync((void**)out, bytes, g_upload_stream));
        keep.push_back(*out);
        if (!v.empty()) {
            CU(cudaMemcpyAsync(*out, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, g_upload_stream));
            CU(cudaStreamSynchronize(g_upload_stream));  // v may be a temporary
        }
    } else {
        CU(cudaMalloc((void**)out, bytes));
        keep.push_back(*out);
        if (!v.empty()) CU(cudaMemcpy(*out, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    }
    return 0;
}

static int build_model_data(const nb200_model_desc& m, int T, NormalModel::Data& d,
                            std::vector<void*>&) {
    d.mu = m.mu;
    d.inv_var = 1.0 / (m.sigma * m.sigma);
    (void)T;
    return 0;
}
static int build_model_data(const nb200_model_desc&, int, FunnelModel::Data& d,
                            std::vector<void*>&) {
    d.unused = 0;
    return 0;
}
static int build_model_data(const nb200_model_desc& m, int T, RadonModel::Data& d,
                            std::vector<void*>& keep) {
    RadonLayout L = build_radon_layout(m.n_obs, m.n_county, m.y, m.county, m.floor, T);
    if (L.G >= 65535) return fail(NB200_EINVAL, "radon: too many observation groups");
    d.J = L.J; d.N = L.N; d.n_steps = L.n_steps; d.G = L.G; d.kmax = L.kmax;
    d.T = T; d.in_smem = 0;
    RadonObs* obs;
    int32_t* group_base;
    uint32_t* group_list;
    int rc;
    if ((rc = to_device(L.obs, &obs, keep))) return rc;
    if ((rc = to_device(L.group_base, &group_base, keep))) return rc;
    if ((rc = to_device(L.group_list, &group_list, keep))) return rc;
    d.obs = obs; d.group_base = group_base; d.group_list = group_list;
    return 0;
}

static int build_model_data(const nb200_model_desc& m, int, CustomModel::Data& d,
                            std::vector<void*>& keep) {
    std::vector<double> v(m.user_data, m.user_data + m.n_user_data);
    double* dev;
    int rc;
    if ((rc = to_device(v, &dev, keep))) return rc;
    d.data = dev;
    d.n_data = (int)m.n_user_data;
    d.n_scratch = (int)m.n_user_scratch;
    d.program = custom_register(m.cuda_source);
    return 0;
}

static int build_model_data(const nb200_model_desc&, int, HostModel::Data& d, std::vector<void*>&) {
    std::memset(&d, 0, sizeof(d));
    return 0;
}

template <class M>
static nb200_sampler* create_impl(const nb200_settings* st, const nb200_model_desc* m,
                                  uint64_t n_chains, uint64_t chain_id_offset, int device,
                                  const double* q0, const double* init_mean) {
    auto* s = new SamplerImpl<M>();
    auto bail = [&](void) -> nb200_sampler* {
        nb200_sampler_destroy(s);
        return nullptr;
    };
    s->device = device;
    s->st = *st;
    s->model = *m;
    s->n_chains = n_chains;
    s->chain_id_offset = chain_id_offset;
    if (cudaSetDevice(device) != cudaSuccess) {
        fail(NB200_ECUDA, "cudaSetDevice failed");
        return bail();
    }
    s->lr = st->adaptation == 1;
    s->W = s->lr ? 1 : pick_W(*m, n_chains);  // low rank: a warp per chain, run-time loops
    if (s->W < 1 || s->W > 32 || (s->W & (s->W - 1))) {
        fail(NB200_EINVAL, "threads per chain must be 32..1024, power of two");
        return bail();
    }
    const int D = (int)m->dim;
    s->Dp = (D + 3) / 4 * 4;
    s->NS = 3 * ((int)st->maxdepth + 1) + 3;
    if (!s->lr) {
        const int forced_t = g_threads_per_chain.load();
        if (s->W == 1 && forced_t < 32 && g_force_nit.load() != 0)
            s->sub = sub_warp_lanes<M>(D, forced_t);
        if (forced_t > 0 && forced_t < 32 && s->sub == 0) {
            fail(NB200_EINVAL, "threads per chain must be 4 (dim <= 12), 8 or 16 (dim <= 2 x threads) for "
                               "this density, or 32..1024 (power of two)");
            return bail();
        }
    }
    if (s->sub) {
        s->NIT = (D + s->sub - 1) / s->sub;
    } else {
        const int T = 32 * s->W;
        s->NIT = supported_nit<M>(s->W, (D + T - 1) / T);
        if (g_force_nit.load() == 0 || s->lr) s->NIT = 0;
        // Two warps per chain (integrator + tree) when the density does enough work per leaf to
        // hide the tree bookkeeping behind it; needs kPipeDepth extra pool slots.
        s->piped = !s->lr && s->W == 1 && s->NIT > 0 && g_pipeline.load() != 0 && D >= 64 &&
                   g_threads_per_chain.load() == 0 && supports_pipeline<M>(s->NIT) &&
                   s->NS + kPipeDepth <= kMaxSlots;
        if (s->piped) s->NS += kPipeDepth;
    }
    s->n_total = st->num_tune + st->num_draws;
    s->n_rows = st->save_warmup ? s->n_total : st->num_draws;
    const bool thinned = st->store_dims && st->store_dims < m->dim;
    const bool expand = st->expand_draws && !thinned;
    s->sdim = thinned ? st->store_dims : (expand ? (uint64_t)M::expanded_dim((int)m->dim) : m->dim);
    s->grad_dim = thinned ? st->store_dims : m->dim;
    KParams<M>& P = s->P;
    std::memset(&P, 0, sizeof(P));
    P.st = *st;
    pool_keep_memory(device);
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking) != cudaSuccess) {
        fail(NB200_ECUDA, "cudaStreamCreate failed");
        return bail();
    }
    g_upload_stream = s->stream;
    const int mrc = build_model_data(*m, 32 * s->W, P.mdata, s->model_allocs);
    g_upload_stream = nullptr;
    s->model_allocs_pooled = true;
    if (mrc != 0) return bail();
    if constexpr (std::is_same<M, HostModel>::value) {
        s->host.reset(new HostService());
        if (s->host->init(*m, n_chains, s->Dp) != 0) return bail();
        HostModel::Data& hd = P.mdata;
        hd.qbox = s->host->d_qbox; hd.gbox = s->host->d_gbox; hd.lpbox = s->host->d_lpbox;
        hd.rcbox = s->host->d_rcbox; hd.req = s->host->d_req; hd.resp = s->host->d_resp;
        hd.stop = s->host->d_stop; hd.chain = 0; hd.Dp = s->Dp;
    }
    // (low rank: every leaf is built on a shared-memory front of four vectors, also for the
    // elementwise densities that stream straight from the pool under the diagonal metric)
    const size_t fixed = smem_for<M>(s->W, P.mdata, s->Dp) +
                         ((s->lr && M::kElementwise) ? align16(4 * sizeof(double) * (size_t)s->Dp) : 0);
    size_t bdata = 0;
    if (s->W == 1) {
        int c = g_chains_per_block.load();
        const size_t bd = model_block_data_bytes<M>(P.mdata);
        if (c <= 0) {
            if (bd > 0) {
                // densities with constant tables: one CTA per SM hosting all of that SM's
                // chains, so the tables are staged in shared memory once per SM
                uint64_t per_sm = (n_chains + 147) / 148;
                c = (int)(per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
            } else {
                c = 1;
                while (c < 8 && (n_chains + c - 1) / c > 148ull * 24) c *= 2;
            }
        }
        if (c > 8) c = 8;
        // (piped: 8 chains = 16 warps = 512 threads, 128 registers per thread)
        if (s->sub) c = 4 * (32 / s->sub);  // four warps per CTA, 32 / lanes chains per warp
        s->cpb = c;
        // stage the tables in shared memory only when they fit beside the chains' own state
        // (a model with > ~10 k observations reads them through L2 instead)
        if (bd > 0 && c >= 4 && bd + (size_t)c * fixed + 2048 <= 227 * 1024) bdata = bd;
    } else {
        s->cpb = 1;
    }
    s->block_data = bdata;
    // Shared-memory budget per chain: the SM's 227 KB divided by the chains that
    // will be co-resident on it (chains spread evenly over the 148 SMs).
    {
        const size_t kSmemSM = 227 * 1024, kSlack = 1024;  // 1 KB/CTA reserved by the driver
        uint64_t per_sm = (n_chains + 147) / 148;
        const int tpc = s->sub ? s->sub : 32 * s->W * (s->piped ? 2 : 1);
        const uint64_t max_res = (uint64_t)(2048 / tpc);  // thread limit per SM
        if (per_sm > max_res) per_sm = max_res;
        if (per_sm > 32ull * s->cpb) per_sm = 32ull * s->cpb;     // CTA limit per SM
        if (per_sm < 1) per_sm = 1;
        const uint64_t ctas = (per_sm + s->cpb - 1) / s->cpb;
        const size_t reserved = ctas * kSlack + ctas * bdata;
        size_t budget = reserved < kSmemSM ? (kSmemSM - reserved) / (ctas * s->cpb) : 0;
        const size_t slot_b = align16(sizeof(double) * 4 * (size_t)s->Dp);
        const size_t var_b = align16(sizeof(double) * (size_t)s->Dp);
        int slots = 0, var_in = 0;
        if (budget > fixed + var_b) {
            var_in = 1;
            slots = (int)((budget - fixed - var_b) / slot_b);
            if (slots > s->NS) slots = s->NS;
        }
        if (s->lr) slots = var_in = 0;  // low rank: the state pool and the metric stay in HBM / L2
        const int forced = g_smem_slots.load();
        if (forced >= 0 && !s->lr) {
            slots = forced > s->NS ? s->NS : forced;
            var_in = (fixed + var_b + slots * slot_b) <= kSmemSM - kSlack;
        }
        P.smem_slots = slots;
        s->smem_slots = slots;
        P.var_in_smem = var_in;
        s->smem_per_chain = chain_smem_total(fixed, s->Dp, var_in, slots);
    }
    s->block = s->sub ? 128 : (s->piped ? 128 * ((s->cpb + 1) / 2) : 32 * s->W * s->cpb);
    s->grid = (int)((n_chains + s->cpb - 1) / s->cpb);
    if (bdata + s->smem_per_chain * s->cpb > 227 * 1024) {
        fail(NB200_EINVAL, "model dimension too large: a density that gathers across dimensions "
                           "keeps 4 vectors of the chain in shared memory (227 KB per SM)");
        return bail();
    }
    if constexpr (std::is_same<M, CustomModel>::value) {
        // compile now so that errors in the user's source surface here, with the NVRTC log
        if (custom_compile(P.mdata.program, s->W, s->lr ? -1 : s->NIT) != 0) {
            fail(NB200_ECOMPILE, custom_last_log());
            return bail();
        }
    }
    P.D = D; P.Dp = s->Dp; P.NS = s->NS;
    P.n_chains = n_chains; P.chain_id_offset = chain_id_offset;
    P.n_rows = s->n_rows; P.sdim = s->sdim; P.n_total = s->n_total;
    P.expand = expand ? 1 : 0;
    P.gdim = s->grad_dim;
    P.stage_loads = g_stage_loads.load();
#define ALLOC(ptr, bytes)                                                               \
    do {                                                                                \
        cudaError_t e_ = cudaMallocAsync((void**)&(ptr), (bytes) ? (bytes) : 8, s->stream); \
        if (e_ != cudaSuccess) {                                                        \
            fail(NB200_ECUDA, std::string("cudaMalloc ") + #ptr + ": " + cudaGetErrorString(e_)); \
            return bail();                                                              \
        }                                                                               \
    } while (0)
#define CHK(call)                                                                       \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            fail(NB200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
            return bail();                                                              \
        }                                                                               \
    } while (0)
    const size_t vecb = sizeof(double) * (size_t)s->Dp;
    ALLOC(s->d_pool, n_chains * (size_t)s->NS * (s->lr ? 5 : 4) * vecb);
    if (s->lr) {
        // window capacity: foreground-only + background + the stretch after the last switch
        const uint64_t f = st->mass_matrix_switch_freq > st->early_mass_matrix_switch_freq
                               ? st->mass_matrix_switch_freq : st->early_mass_matrix_switch_freq;
        s->lr_cap = (int)(3 * f + 2 < s->n_total + 2 ? 3 * f + 2 : s->n_total + 2);
        s->lr_max_rank = (int)(st->mass_matrix_max_rank < m->dim ? st->mass_matrix_max_rank : m->dim);
        const size_t R = (size_t)s->lr_max_rank;
        ALLOC(s->d_lr_stds, n_chains * vecb);
        ALLOC(s->d_lr_vals, n_chains * R * sizeof(double));
        ALLOC(s->d_lr_vecs, n_chains * R * vecb);
        ALLOC(s->d_lr_coef, n_chains * R * sizeof(double));
        ALLOC(s->d_lr_win, n_chains * (size_t)s->lr_cap * 2 * vecb);
        ALLOC(s->d_lr_mat, n_chains * 2 * (size_t)D * vecb);
        ALLOC(s->d_lr_cols, n_chains * 6 * vecb);
        if (st->store_mass_matrix) ALLOC(s->d_eig, n_chains * s->n_rows * R * sizeof(double));
    }
    ALLOC(s->d_var, n_chains * vecb);
    ALLOC(s->d_wf, n_chains * 8 * vecb);
    ALLOC(s->d_sc, n_chains * sizeof(ChainScalars));
    ALLOC(s->d_draws, n_chains * s->n_rows * s->sdim * sizeof(double));
    ALLOC(s->d_stats, n_chains * s->n_rows * NB200_NSTAT * sizeof(double));
    if (st->store_gradient) ALLOC(s->d_grads, n_chains * s->n_rows * s->grad_dim * sizeof(double));
    if (st->store_mass_matrix) ALLOC(s->d_mm, n_chains * s->n_rows * s->grad_dim * sizeof(double));
    if (st->store_divergences) {
        const size_t nb = n_chains * s->n_rows * 4 * s->grad_dim * sizeof(double);
        ALLOC(s->d_div, nb);
        CHK(cudaMemsetAsync(s->d_div, 0xFF, nb, s->stream));  // all-ones = NaN: "did not diverge"
    }
    ALLOC(s->d_stop, sizeof(int));
    CHK(cudaEventCreate(&s->ev0));
    CHK(cudaEventCreate(&s->ev1));
    CHK(cudaMemsetAsync(s->d_sc, 0, n_chains * sizeof(ChainScalars), s->stream));
    CHK(cudaMemsetAsync(s->d_stop, 0, sizeof(int), s->stream));
    s->h_sc_store.resize(n_chains);
    s->h_sc = s->h_sc_store.data();
    std::memset(s->h_sc, 0, n_chains * sizeof(ChainScalars));
    if (q0) {
        ALLOC(s->d_q0, n_chains * D * sizeof(double));
        CHK(cudaMemcpyAsync(s->d_q0, q0, n_chains * D * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    }
    if (init_mean) {
        ALLOC(s->d_init_mean, D * sizeof(double));
        CHK(cudaMemcpyAsync(s->d_init_mean, init_mean, D * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    }
    CHK(cudaStreamSynchronize(s->stream));
#undef ALLOC
#undef CHK
    P.pool = s->d_pool; P.var = s->d_var; P.welford = s->d_wf; P.sc = s->d_sc;
    P.draws = s->d_draws; P.stats = s->d_stats; P.grads = s->d_grads; P.mminv = s->d_mm;
    P.q0 = s->d_q0; P.init_mean = s->d_init_mean; P.z_tape = nullptr;
    P.divs = s->d_div;
    P.lr_stds = s->d_lr_stds; P.lr_vals = s->d_lr_vals; P.lr_vecs = s->d_lr_vecs;
    P.lr_coef = s->d_lr_coef; P.lr_win = s->d_lr_win; P.lr_mat = s->d_lr_mat;
    P.lr_cols = s->d_lr_cols; P.eigvals = s->d_eig;
    P.lr_cap = s->lr_cap; P.lr_max_rank = s->lr_max_rank;
    // host plug-in: the stop flag lives in mapped pinned memory so that the service thread that
    // meets a fatal return code can raise it without a CUDA call
    P.stop_flag = s->host ? s->host->d_stop : s->d_stop;
    s->rows_filled.assign(n_chains, 0);
    // Streaming regime (a CTA per chain, state pool >> L2): the mass matrix is the one vector that
    // EVERY pass of a chain reads again (leapfrog and U-turn passes alike: 1 of 7 and 1 of 5
    // vectors moved).  All chains' copies together are a few tens of MB — they fit the L2 if the
    // terabytes streaming past do not evict them: an access-policy window marks them persisting in
    // a set-aside part of the L2 (NB200_L2_PERSIST=0 turns it off for A/B runs).
    if (M::kElementwise && s->W >= 4 && s->smem_slots == 0) {
        const char* env = std::getenv("NB200_L2_PERSIST");
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device);
        const size_t bytes = n_chains * sizeof(double) * (size_t)s->Dp;
        if (!(env && env[0] == '0') && max_persist > 0 && max_window > 0) {
            const size_t carve = bytes < (size_t)max_persist ? bytes : (size_t)max_persist;
            const size_t window = bytes < (size_t)max_window ? bytes : (size_t)max_window;
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
                cudaStreamAttrValue attr;
                std::memset(&attr, 0, sizeof(attr));
                attr.accessPolicyWindow.base_ptr = s->d_var;
                attr.accessPolicyWindow.num_bytes = window;
                attr.accessPolicyWindow.hitRatio = window > 0 ? (float)((double)carve / (double)window) : 0.f;
                if (attr.accessPolicyWindow.hitRatio > 1.f) attr.accessPolicyWindow.hitRatio = 1.f;
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                if (cudaStreamSetAttribute(s->stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess) {
                    s->l2_persist = true;
                    g_persist_users[device & 15].fetch_add(1);
                }
            }
            cudaGetLastError();  // the window is an optimisation: never an error
        }
    }
    return s;
}

// ------------------------------------------------------------------ C-ABI
extern "C" {

int nb200_abi_version(void) { return NB200_ABI_VERSION; }
const char* nb200_last_error(void) { return g_err.c_str(); }
uint64_t nb200_model_expanded_dim(const nb200_model_desc* model) {
    if (!model) return 0;
    switch (model->kind) {
    case NB200_MODEL_NORMAL: return (uint64_t)NormalModel::expanded_dim((int)model->dim);
    case NB200_MODEL_FUNNEL: return (uint64_t)FunnelModel::expanded_dim((int)model->dim);
    case NB200_MODEL_RADON: return (uint64_t)RadonModel::expanded_dim((int)model->dim);
    case NB200_MODEL_CUSTOM: return (uint64_t)CustomModel::expanded_dim((int)model->dim);
    case NB200_MODEL_HOST: return (uint64_t)HostModel::expanded_dim((int)model->dim);
    }
    return 0;
}
int nb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
void nb200_set_threads_per_chain(int32_t t) { g_threads_per_chain.store(t); }
void nb200_set_chains_per_block(int32_t c) { g_chains_per_block.store(c); }
void nb200_set_smem_slots(int32_t n) { g_smem_slots.store(n); }
void nb200_set_unroll(int32_t on) { g_force_nit.store(on ? -1 : 0); }
void nb200_set_stage_loads(int32_t mode) { g_stage_loads.store(mode & 7); }
void nb200_set_pipeline(int32_t on) { g_pipeline.store(on ? 1 : 0); }

void nb200_settings_default(nb200_settings* s) {
    std::memset(s, 0, sizeof(*s));
    s->seed = 0;
    s->num_tune = 400;
    s->num_draws = 1000;
    s->maxdepth = 10;
    s->mindepth = 0;
    s->check_turning = 1;
    s->store_gradient = 0;
    s->store_mass_matrix = 0;
    s->use_grad_based_estimate = 1;
    s->max_energy_error = 1000.0;
    s->initial_step = 0.1;
    s->target_accept = 0.8;
    s->max_step_size = INFINITY;
    s->da_k = 0.75;
    s->da_t0 = 10.0;
    s->da_gamma = 0.05;
    s->step_size_method = 0;
    s->fixed_step_size = 0.1;
    s->early_window = 0.3;
    s->step_size_window = 0.15;
    s->mass_matrix_switch_freq = 80;
    s->early_mass_matrix_switch_freq = 10;
    s->mass_matrix_update_freq = 1;
    s->init_kind = 0;
    s->num_try_init = 10;
    s->init_radius = 2.0;
    s->store_dims = 0;
    s->save_warmup = 1;
    s->expand_draws = 0;
    s->store_divergences = 0;
    s->adaptation = 0;
    s->adam_learning_rate = 0.05;
    s->step_size_jitter = 0.0;
    s->mass_matrix_eigval_cutoff = 2.0;
    s->mass_matrix_gamma = 1e-5;
    s->mass_matrix_max_rank = 32;
}

void* nb200_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 8, cudaHostAllocDefault) != cudaSuccess) {
        fail(NB200_ECUDA, "cudaHostAlloc failed");
        return nullptr;
    }
    return p;
}
void nb200_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

nb200_sampler* nb200_sampler_create(const nb200_settings* settings, const nb200_model_desc* model,
                                    uint64_t n_chains, uint64_t chain_id_offset, int device,
                                    const double* q0, const double* init_mean) {
    if (validate(settings, model) != 0) return nullptr;
    if (n_chains < 1) {
        fail(NB200_EINVAL, "n_chains must be >= 1");
        return nullptr;
    }
    if (chain_id_offset + n_chains > 0xFFFFFFFFull) {
        fail(NB200_EINVAL, "global chain id must fit 32 bits");
        return nullptr;
    }
    if (settings->num_tune + settings->num_draws >= 0xFFFFFFFFull) {
        fail(NB200_EINVAL, "num_tune + num_draws must fit 32 bits");
        return nullptr;
    }
    int ndev = nb200_device_count();
    if (ndev < 1) {
        fail(NB200_ECUDA, "no CUDA device available: the B200 engine has no CPU fallback");
        return nullptr;
    }
    if (device < 0 || device >= ndev) {
        fail(NB200_EINVAL, "device index out of range");
        return nullptr;
    }
    switch (model->kind) {
    case NB200_MODEL_NORMAL:
        return create_impl<NormalModel>(settings, model, n_chains, chain_id_offset, device, q0, init_mean);
    case NB200_MODEL_FUNNEL:
        return create_impl<FunnelModel>(settings, model, n_chains, chain_id_offset, device, q0, init_mean);
    case NB200_MODEL_RADON:
        return create_impl<RadonModel>(settings, model, n_chains, chain_id_offset, device, q0, init_mean);
    case NB200_MODEL_CUSTOM:
        return create_impl<CustomModel>(settings, model, n_chains, chain_id_offset, device, q0, init_mean);
    case NB200_MODEL_HOST:
        return create_impl<HostModel>(settings, model, n_chains, chain_id_offset, device, q0, init_mean);
    }
    fail(NB200_EINVAL, "unknown model kind");
    return nullptr;
}

int nb200_sampler_trace_bytes(nb200_sampler* s, size_t* draws_bytes, size_t* stats_bytes) {
    if (!s) return fail(NB200_EINVAL, "null sampler");
    if (draws_bytes) *draws_bytes = s->n_chains * s->n_rows * s->sdim * sizeof(double);
    if (stats_bytes) *stats_bytes = s->n_chains * s->n_rows * NB200_NSTAT * sizeof(double);
    return 0;
}

int nb200_sampler_set_trace_target(nb200_sampler* s, double* draws, size_t draws_bytes, double* stats,
                                   size_t stats_bytes) {
    if (!s) return fail(NB200_EINVAL, "null sampler");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state != RunState::Created) return fail(NB200_ESTATE, "trace target must be set before start");
    // rows are streamed into these buffers while the kernel runs: a wrong size is a host
    // memory overrun, so it is refused here and not discovered after sampling
    if (draws && draws_bytes != s->n_chains * s->n_rows * s->sdim * sizeof(double))
        return fail(NB200_EINVAL, "trace target: draws buffer must hold n_rows * n_chains * width doubles");
    if (stats && stats_bytes != s->n_chains * s->n_rows * NB200_NSTAT * sizeof(double))
        return fail(NB200_EINVAL, "trace target: stats buffer must hold n_rows * n_chains * 16 doubles");
    s->tgt_draws = draws;
    s->tgt_stats = stats;
    // pageable targets: 2 MB pages, so that landing the rows is not a page fault per 4 KB
    auto advise = [](void* ptr, size_t bytes) {
        if (!ptr || bytes < (4u << 20) || is_pinned_host(ptr)) return;
        const uintptr_t lo = ((uintptr_t)ptr + 4095) & ~uintptr_t(4095);
        const uintptr_t hi = ((uintptr_t)ptr + bytes) & ~uintptr_t(4095);
        if (hi > lo) madvise((void*)lo, hi - lo, MADV_HUGEPAGE);
    };
    advise(draws, draws_bytes);
    advise(stats, stats_bytes);
    return 0;
}

int nb200_sampler_set_trace_target_strided(nb200_sampler* s, double* draws, size_t draws_row_stride,
                                           double* stats, size_t stats_row_stride) {
    if (!s) return fail(NB200_EINVAL, "null sampler");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state != RunState::Created) return fail(NB200_ESTATE, "trace target must be set before start");
    if (draws && draws_row_stride < s->n_chains * s->sdim)
        return fail(NB200_EINVAL, "trace target: draws row stride must be >= n_chains * width doubles");
    if (stats && stats_row_stride < s->n_chains * NB200_NSTAT)
        return fail(NB200_EINVAL, "trace target: stats row stride must be >= n_chains * 16 doubles");
    s->tgt_draws = draws;
    s->tgt_stats = stats;
    s->tgt_draws_pitch = draws ? draws_row_stride : 0;
    s->tgt_stats_pitch = stats ? stats_row_stride : 0;
    return 0;
}

int nb200_sampler_set_draws_per_launch(nb200_sampler* s, uint64_t n) {
    if (!s) return fail(NB200_EINVAL, "null sampler");
    std::lock_guard<std::mutex> lk(s->mu);
    s->draws_per_launch = n;
    return 0;
}

}  // extern "C"
// set the tape pointer inside the model-typed KParams
template <class M>
static void set_tape(nb200_sampler* s) {
    static_cast<SamplerImpl<M>*>(s)->P.z_tape = s->d_tape;
}
extern "C" {

int nb200_sampler_set_z_tape(nb200_sampler* s, const double* z_tape) {
    if (!s || !z_tape) return fail(NB200_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state != RunState::Created) return fail(NB200_ESTATE, "tape must be set before start");
    CU(cudaSetDevice(s->device));
    const size_t n = s->n_chains * s->n_total * s->model.dim;
    CU(cudaMalloc((void**)&s->d_tape, n * sizeof(double)));
    CU(cudaMemcpy(s->d_tape, z_tape, n * sizeof(double), cudaMemcpyHostToDevice));
    switch (s->model.kind) {
    case NB200_MODEL_NORMAL: set_tape<NormalModel>(s); break;
    case NB200_MODEL_FUNNEL: set_tape<FunnelModel>(s); break;
    case NB200_MODEL_RADON: set_tape<RadonModel>(s); break;
    case NB200_MODEL_CUSTOM: set_tape<CustomModel>(s); break;
    case NB200_MODEL_HOST: set_tape<HostModel>(s); break;
    }
    return 0;
}

static int launch_locked(nb200_sampler* s) {
    CU(cudaSetDevice(s->device));
    CU(cudaEventRecord(s->ev0, s->stream));
    int rc = s->launch();
    if (rc != 0) return rc;
    CU(cudaEventRecord(s->ev1, s->stream));
    s->launches += 1;
    return 0;
}

int nb200_sampler_start(nb200_sampler* s) {
    if (!s) return fail(NB200_EINVAL, "null sampler");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state != RunState::Created) return fail(NB200_ESTATE, "sampler already started");
    int rc = launch_locked(s);
    if (rc != 0) {
        s->state = RunState::Error;
        s->sampler_error = rc;
        s->err = g_err;
        return rc;
    }
    s->state = RunState::Running;
    return 0;
}

// ---- device -> host copies of trace blocks ---------------------------------------------------
// A PINNED destination takes the DMA directly.  A pageable one (the arrays a plain
// nutpie_b200.sample() call returns) would go through the driver's single-threaded bounce copy
// and fault its pages in one by one (measured 1.6 s for the 5.9 GB trace of the BASELINE radon
// job, 7x the sampling itself): instead the block is cut into chunks that land in a process-wide
// ring of two pinned 64 MB buffers and are moved on by a few host threads while the next
// chunk's DMA runs, with the destination advised to use huge pages.
struct StageRing {
    std::mutex mu;
    char* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    static constexpr size_t kChunk = 64ull << 20;
    int init() {
        if (buf[0]) return 0;
        for (int i = 0; i < 2; ++i) {
            if (cudaHostAlloc((void**)&buf[i], kChunk, cudaHostAllocPortable) != cudaSuccess) return -1;
            if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) return -1;
        }
        return 0;
    }
};
static StageRing g_stage_of_device[16];  // (one per device: the shards of a multi-GPU job stream concurrently)

static bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

static void parallel_memcpy(char* dst, const char* src, size_t bytes, int T) {
    if (T <= 1 || bytes < (4u << 20)) {
        std::memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = ((bytes / T) + 4095) & ~size_t(4095);
    for (int t = 0; t < T; ++t) {
        const size_t lo = per * t, hi = lo + per < bytes ? lo + per : bytes;
        if (lo >= bytes) break;
        th.emplace_back([=] { std::memcpy(dst + lo, src + lo, hi - lo); });
    }
    for (auto& t : th) t.join();
}

// copy `bytes` from device memory to host memory on `stream`; returns when the data has landed
static int d2h_block(void* dst, const void* src, size_t bytes, cudaStream_t stream, int device) {
    StageRing& g_stage = g_stage_of_device[device & 15];
    if (bytes == 0) return 0;
    if (is_pinned_host(dst) || bytes < (8u << 20)) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        return 0;
    }
    std::lock_guard<std::mutex> lk(g_stage.mu);
    if (g_stage.init() != 0) {  // no staging memory: the driver's own path
        cudaGetLastError();
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        return 0;
    }
    int T = usable_cores() / 2;
    T = T < 1 ? 1 : (T > 8 ? 8 : T);
    const size_t CH = StageRing::kChunk;
    const size_t n = (bytes + CH - 1) / CH;
    auto len = [&](size_t c) { return c + 1 < n ? CH : bytes - c * CH; };
    CU(cudaMemcpyAsync(g_stage.buf[0], src, len(0), cudaMemcpyDeviceToHost, stream));
    CU(cudaEventRecord(g_stage.ev[0], stream));
    for (size_t c = 0; c < n; ++c) {
        const int b = (int)(c & 1);
        if (c + 1 < n) {  // next chunk's DMA runs while this one is moved on
            CU(cudaMemcpyAsync(g_stage.buf[b ^ 1], (const char*)src + (c + 1) * CH, len(c + 1),
                               cudaMemcpyDeviceToHost, stream));
            CU(cudaEventRecord(g_stage.ev[b ^ 1], stream));
        }
        CU(cudaEventSynchronize(g_stage.ev[b]));
        parallel_memcpy((char*)dst + c * CH, g_stage.buf[b], len(c), T);
    }
    return 0;
}

// `n_rows` rows of `row` doubles, dense on the device, to a host buffer whose rows are `pitch`
// doubles apart (a shard of a multi-GPU job writes its chains into ITS columns of the job's one
// [row][chain][width] array): dense -> d2h_block; pinned -> one 2-D DMA; pageable -> whole rows
// through the staging ring, moved on row by row by the copy threads.
static int d2h_rows(double* dst, size_t pitch, const double* src, size_t row, size_t n_rows,
                    cudaStream_t stream, int device) {
    if (n_rows == 0 || row == 0) return 0;
    if (pitch == 0 || pitch == row) return d2h_block(dst, src, n_rows * row * sizeof(double), stream, device);
    if (is_pinned_host(dst)) {
        CU(cudaMemcpy2DAsync(dst, pitch * sizeof(double), src, row * sizeof(double), row * sizeof(double),
                             n_rows, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        return 0;
    }
    StageRing& g_stage = g_stage_of_device[device & 15];
    std::lock_guard<std::mutex> lk(g_stage.mu);
    const size_t rb = row * sizeof(double);
    if (g_stage.init() != 0 || rb > StageRing::kChunk) {  // no staging: the driver's own 2-D path
        cudaGetLastError();
        CU(cudaMemcpy2DAsync(dst, pitch * sizeof(double), src, rb, rb, n_rows, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        return 0;
    }
    int T = usable_cores() / 2;
    T = T < 1 ? 1 : (T > 8 ? 8 : T);
    const size_t per = StageRing::kChunk / rb;  // rows per chunk
    const size_t n = (n_rows + per - 1) / per;
    auto rows_of = [&](size_t c) { return c + 1 < n ? per : n_rows - c * per; };
    CU(cudaMemcpyAsync(g_stage.buf[0], src, rows_of(0) * rb, cudaMemcpyDeviceToHost, stream));
    CU(cudaEventRecord(g_stage.ev[0], stream));
    for (size_t c = 0; c < n; ++c) {
        const int b = (int)(c & 1);
        if (c + 1 < n) {
            CU(cudaMemcpyAsync(g_stage.buf[b ^ 1], src + (c + 1) * per * row, rows_of(c + 1) * rb,
                               cudaMemcpyDeviceToHost, stream));
            CU(cudaEventRecord(g_stage.ev[b ^ 1], stream));
        }
        CU(cudaEventSynchronize(g_stage.ev[b]));
        const size_t nr = rows_of(c);
        double* d0 = dst + c * per * pitch;
        const char* s0 = g_stage.buf[b];
        std::vector<std::thread> th;
        const int TT = (size_t)T < nr ? T : (int)nr;
        for (int t = 0; t < TT; ++t) {
            const size_t lo = nr * t / TT, hi = nr * (t + 1) / TT;
            th.emplace_back([=] {
                for (size_t r = lo; r < hi; ++r) std::memcpy(d0 + r * pitch, s0 + r * rb, rb);
            });
        }
        for (auto& t : th) t.join();
    }
    return 0;
}

// refresh h_sc from the device on the side stream (safe while the kernel runs)
static int fetch_scalars(nb200_sampler* s);
static int stream_rows(nb200_sampler* s, uint64_t from, uint64_t to);
static int fetch_scalars(nb200_sampler* s) {
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(s->h_sc, s->d_sc, s->n_chains * sizeof(ChainScalars), cudaMemcpyDeviceToHost,
                       s->side));
    CU(cudaStreamSynchronize(s->side));
    return 0;
}

// rows [from, to) of every chain -> the registered host buffers: with the [row][chain][...]
// layout this is one contiguous block per buffer
static int stream_rows(nb200_sampler* s, uint64_t from, uint64_t to) {
    if (to <= from) return 0;
    CU(cudaSetDevice(s->device));
    const size_t drow = s->n_chains * s->sdim, srow = s->n_chains * NB200_NSTAT;  // doubles per row
    int rc = 0;
    const size_t dp = s->tgt_draws_pitch ? s->tgt_draws_pitch : drow, sp = s->tgt_stats_pitch ? s->tgt_stats_pitch : srow;
    if (s->tgt_draws)
        rc = d2h_rows(s->tgt_draws + from * dp, dp, s->d_draws + from * drow, drow, to - from, s->side, s->device);
    if (rc == 0 && s->tgt_stats)
        rc = d2h_rows(s->tgt_stats + from * sp, sp, s->d_stats + from * srow, srow, to - from, s->side, s->device);
    if (rc != 0) return rc;
    s->streamed_rows = to;
    return 0;
}

// while the kernel runs: copy out the rows every chain has already published
static int stream_progress(nb200_sampler* s) {
    if (!s->tgt_draws && !s->tgt_stats) return 0;
    auto now = std::chrono::steady_clock::now();
    if (std::chrono::duration<double>(now - s->last_stream).count() < 0.004) return 0;
    s->last_stream = now;
    int rc = fetch_scalars(s);
    if (rc != 0) return rc;
    uint64_t m = ~0ull;
    for (uint64_t c = 0; c < s->n_chains; ++c)
        if (s->h_sc[c].published < m) m = s->h_sc[c].published;
    uint64_t rows = s->st.save_warmup ? m : (m > s->st.num_tune ? m - s->st.num_tune : 0);
    if (rows > s->n_rows) rows = s->n_rows;
    if (rows >= s->streamed_rows + 32) return stream_rows(s, s->streamed_rows, rows);
    return 0;
}

// called with the lock held when the current launch has completed
static int on_launch_done(nb200_sampler* s) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    s->kernel_ms += ms;
    int rc = fetch_scalars(s);
    if (rc != 0) return rc;
    bool all_done = true;
    for (uint64_t c = 0; c < s->n_chains; ++c) {
        if (s->h_sc[c].status < 0) {
            s->state = RunState::Error;
            s->sampler_error = s->h_sc[c].status;
            s->err = s->h_sc[c].status == NB200_EINIT
                         ? "chain " + std::to_string(c) + ": no finite initial point found"
                         : "chain " + std::to_string(c) + ": fatal logp error";
            return fail(s->h_sc[c].status, s->err);
        }
        if (s->h_sc[c].status != 2) all_done = false;
    }
    if (s->host && s->host->fatal_rc.load() != 0) {
        // src/pymc.rs:166-181: a negative return code is not recoverable — the sampler stops,
        // wait() raises, the draws finished so far stay readable (trace / trace_into)
        s->state = RunState::Error;
        s->sampler_error = NB200_ELOGP;
        s->err = "Logp function returned error code: " + std::to_string(s->host->fatal_rc.load()) +
                 " (chain " + std::to_string(s->host->fatal_chain.load()) + ")";
        return fail(NB200_ELOGP, s->err);
    }
    if (all_done) {
        if (s->tgt_draws || s->tgt_stats) {
            rc = stream_rows(s, s->streamed_rows, s->n_rows);
            if (rc != 0) return rc;
        }
        s->state = RunState::Finished;
        return 0;
    }
    return 1;  // more draws to do
}

int nb200_sampler_wait(nb200_sampler* s, double timeout_seconds) {
    if (!s) return fail(NB200_EINVAL, "null sampler");
    auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        {
            std::lock_guard<std::mutex> lk(s->mu);
            switch (s->state) {
            case RunState::Created: return fail(NB200_ESTATE, "sampler not started");
            case RunState::Finished:
            case RunState::Aborted: return NB200_OK;
            case RunState::Error: return fail(s->sampler_error, s->err.empty() ? "sampler error" : s->err);
            case RunState::Paused: break;
            case RunState::Running: {
                CU(cudaSetDevice(s->device));
                cudaError_t q = cudaEventQuery(s->ev1);
                if (q == cudaSuccess) {
                    int rc = on_launch_done(s);
                    if (rc < 0) return rc;
                    if (rc == 0) return NB200_OK;
                    rc = launch_locked(s);  // chunked mode: next launch
                    if (rc != 0) {
                        s->state = RunState::Error;
                        s->sampler_error = rc;
                        s->err = g_err;
                        return rc;
                    }
                } else if (q == cudaErrorNotReady) {
                    int rc = stream_progress(s);
                    if (rc < 0) return rc;
                } else {
                    s->state = RunState::Error;
                    s->sampler_error = NB200_ECUDA;
                    s->err = std::string("kernel failed: ") + cudaGetErrorString(q);
                    return fail(NB200_ECUDA, s->err);
                }
                break;
            }
            }
        }
        if (timeout_seconds >= 0) {
            double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (el >= timeout_seconds) return NB200_ETIMEOUT;
        }
        std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
}

int nb200_sampler_is_finished(nb200_sampler* s) {
    if (!s) return 0;
    int rc = nb200_sampler_wait(s, 0.0);
    std::lock_guard<std::mutex> lk(s->mu);
    (void)rc;
    return s->state == RunState::Finished || s->state == RunState::Aborted ||
           s->state == RunState::Error;
}

int nb200_sampler_progress(nb200_sampler* s, nb200_progress* out) {
    if (!s || !out) return fail(NB200_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    int rc = fetch_scalars(s);
    if (rc != 0) return rc;
    for (uint64_t c = 0; c < s->n_chains; ++c) {
        const ChainScalars& sc = s->h_sc[c];
        out[c].finished_draws = sc.draw;
        out[c].total_draws = s->n_total;
        out[c].divergences = sc.divergences;
        out[c].latest_num_steps = sc.latest_n_steps;
        out[c].total_num_steps = sc.total_steps;
        out[c].step_size = sc.step_size;
        out[c].tuning = sc.draw < s->st.num_tune;
        out[c].started = sc.status != 0;
    }
    return 0;
}

static int stop_and_drain(nb200_sampler* s) {
    CU(cudaSetDevice(s->device));
    int one = 1;
    if (s->host) {  // mapped flag; the service threads keep answering until the kernel has left
        __atomic_store_n(s->host->stop, 1, __ATOMIC_RELEASE);
    } else {
        CU(cudaMemcpyAsync(s->d_stop, &one, sizeof(int), cudaMemcpyHostToDevice, s->side));
        CU(cudaStreamSynchronize(s->side));
    }
    CU(cudaStreamSynchronize(s->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    s->kernel_ms += ms;
    int zero = 0;
    if (s->host) __atomic_store_n(s->host->stop, 0, __ATOMIC_RELEASE);
    else CU(cudaMemcpy(s->d_stop, &zero, sizeof(int), cudaMemcpyHostToDevice));
    return fetch_scalars(s);
}

int nb200_sampler_pause(nb200_sampler* s) {
    if (!s) return fail(NB200_EINVAL, "null sampler");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state != RunState::Running) return 0;
    int rc = stop_and_drain(s);
    if (rc != 0) return rc;
    bool all_done = true;
    for (uint64_t c = 0; c < s->n_chains; ++c)
        if (s->h_sc[c].status != 2) all_done = false;
    s->state = all_done ? RunState::Finished : RunState::Paused;
    return 0;
}

int nb200_sampler_resume(nb200_sampler* s) {
    if (!s) return fail(NB200_EINVAL, "null sampler");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state != RunState::Paused) return 0;
    int rc = launch_locked(s);
    if (rc != 0) return rc;
    s->state = RunState::Running;
    return 0;
}

int nb200_sampler_abort(nb200_sampler* s) {
    if (!s) return fail(NB200_EINVAL, "null sampler");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state == RunState::Running) {
        int rc = stop_and_drain(s);
        if (rc != 0) return rc;
    }
    if (s->state == RunState::Running || s->state == RunState::Paused) s->state = RunState::Aborted;
    return 0;
}

static int copy_trace(nb200_sampler* s, double* draws, double* stats, double* grads, double* mm,
                      uint64_t* rows) {
    CU(cudaSetDevice(s->device));
    int rc = fetch_scalars(s);
    if (rc != 0) return rc;
    const bool live = s->state == RunState::Running;
    for (uint64_t c = 0; c < s->n_chains; ++c) {
        // while the kernel runs only the rows it has fenced and published are safe to read
        uint64_t d = live ? s->h_sc[c].published : s->h_sc[c].draw;
        uint64_t r = s->st.save_warmup ? d : (d > s->st.num_tune ? d - s->st.num_tune : 0);
        s->rows_filled[c] = r < s->n_rows ? r : s->n_rows;
        if (rows) rows[c] = s->rows_filled[c];
    }
    const size_t nd = s->n_chains * s->n_rows * s->sdim * sizeof(double);
    const size_t ns = s->n_chains * s->n_rows * NB200_NSTAT * sizeof(double);
    const bool streamed = s->streamed_rows >= s->n_rows;  // already landed in the target buffers
    // (the registered targets may be row-strided: a caller that passes them back gets their pitch)
    const size_t drow = s->n_chains * s->sdim, srow = s->n_chains * NB200_NSTAT;
    if (draws && !(streamed && draws == s->tgt_draws) &&
        (rc = d2h_rows(draws, draws == s->tgt_draws ? s->tgt_draws_pitch : 0, s->d_draws, drow, s->n_rows, s->side, s->device)))
        return rc;
    if (stats && !(streamed && stats == s->tgt_stats) &&
        (rc = d2h_rows(stats, stats == s->tgt_stats ? s->tgt_stats_pitch : 0, s->d_stats, srow, s->n_rows, s->side, s->device)))
        return rc;
    (void)nd; (void)ns;
    const size_t ng = s->n_chains * s->n_rows * s->grad_dim * sizeof(double);
    if (grads && s->d_grads && (rc = d2h_block(grads, s->d_grads, ng, s->side, s->device))) return rc;
    if (mm && s->d_mm && (rc = d2h_block(mm, s->d_mm, ng, s->side, s->device))) return rc;
    return 0;
}

int nb200_sampler_trace_into(nb200_sampler* s, double* draws, double* stats, double* gradients,
                             double* mass_matrix_inv, uint64_t* rows_filled) {
    if (!s) return fail(NB200_EINVAL, "null sampler");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state == RunState::Created) return fail(NB200_ESTATE, "sampler not started");
    return copy_trace(s, draws, stats, gradients, mass_matrix_inv, rows_filled);
}

int nb200_sampler_trace(nb200_sampler* s, nb200_trace_view* out) {
    if (!s || !out) return fail(NB200_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state == RunState::Created) return fail(NB200_ESTATE, "sampler not started");
    CU(cudaSetDevice(s->device));
    const size_t nd = s->n_chains * s->n_rows * s->sdim * sizeof(double);
    const size_t ns = s->n_chains * s->n_rows * NB200_NSTAT * sizeof(double);
    if (!s->h_draws) CU(cudaHostAlloc((void**)&s->h_draws, nd ? nd : 8, cudaHostAllocDefault));
    if (!s->h_stats) CU(cudaHostAlloc((void**)&s->h_stats, ns ? ns : 8, cudaHostAllocDefault));
    const size_t ng = s->n_chains * s->n_rows * s->grad_dim * sizeof(double);
    if (s->d_grads && !s->h_grads) CU(cudaHostAlloc((void**)&s->h_grads, ng ? ng : 8, cudaHostAllocDefault));
    if (s->d_mm && !s->h_mm) CU(cudaHostAlloc((void**)&s->h_mm, ng ? ng : 8, cudaHostAllocDefault));
    int rc = copy_trace(s, s->h_draws, s->h_stats, s->h_grads, s->h_mm, nullptr);
    if (rc != 0) return rc;
    out->n_chains = s->n_chains;
    out->n_rows = s->n_rows;
    out->dim = s->model.dim;
    out->store_dims = s->sdim;
    out->draws = s->h_draws;
    out->stats = s->h_stats;
    out->gradients = s->h_grads;
    out->mass_matrix_inv = s->h_mm;
    out->rows_filled = s->rows_filled.data();
    return 0;
}

int nb200_sampler_divergence_trace_into(nb200_sampler* s, double* divergences) {
    if (!s || !divergences) return fail(NB200_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->state == RunState::Created) return fail(NB200_ESTATE, "sampler not started");
    if (!s->d_div) return fail(NB200_EINVAL, "store_divergences was not set");
    CU(cudaSetDevice(s->device));
    const size_t nb = s->n_chains * s->n_rows * 4 * s->grad_dim * sizeof(double);
    CU(cudaMemcpyAsync(divergences, s->d_div, nb, cudaMemcpyDeviceToHost, s->side));
    CU(cudaStreamSynchronize(s->side));
    return 0;
}

int nb200_sampler_eigvals_trace_into(nb200_sampler* s, double* eigvals, uint64_t* max_rank) {
    if (!s) return fail(NB200_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (max_rank) *max_rank = (uint64_t)s->lr_max_rank;
    if (!eigvals) return 0;  // size query
    if (s->state == RunState::Created) return fail(NB200_ESTATE, "sampler not started");
    if (!s->d_eig) return fail(NB200_EINVAL, "needs adaptation = low_rank and store_mass_matrix");
    CU(cudaSetDevice(s->device));
    const size_t nb = s->n_chains * s->n_rows * (size_t)s->lr_max_rank * sizeof(double);
    CU(cudaMemcpyAsync(eigvals, s->d_eig, nb, cudaMemcpyDeviceToHost, s->side));
    CU(cudaStreamSynchronize(s->side));
    return 0;
}

int nb200_host_expand_rows(nb200_expand_fn fn, const void* user_data, size_t dim, size_t expanded_dim,
                           uint64_t n, const double* q, size_t q_stride, double* out, int n_threads) {
    if (!fn || !q || !out) return fail(NB200_EINVAL, "null argument");
    int T = n_threads > 0 ? n_threads : usable_cores();
    if ((uint64_t)T > n) T = n ? (int)n : 1;
    std::atomic<int> bad{0};
    auto work = [&](int t) {
        const uint64_t lo = n * (uint64_t)t / T, hi = n * (uint64_t)(t + 1) / T;
        for (uint64_t i = lo; i < hi && bad.load(std::memory_order_relaxed) == 0; ++i) {
            const int rc = fn(dim, expanded_dim, q + i * q_stride, out + i * expanded_dim, user_data);
            if (rc != 0) {
                int expect = 0;
                bad.compare_exchange_strong(expect, rc);
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    if (bad.load() != 0)
        return fail(NB200_ELOGP, "Expand function returned error code " + std::to_string(bad.load()));
    return 0;
}

double nb200_sampler_kernel_ms(nb200_sampler* s) {
    if (!s) return 0.0;
    std::lock_guard<std::mutex> lk(s->mu);
    return s->kernel_ms;
}
uint64_t nb200_sampler_launch_count(nb200_sampler* s) {
    if (!s) return 0;
    std::lock_guard<std::mutex> lk(s->mu);
    return s->launches;
}
int nb200_sampler_device_buffers(nb200_sampler* s, void** draws, void** stats) {
    if (!s) return fail(NB200_EINVAL, "null sampler");
    if (draws) *draws = s->d_draws;
    if (stats) *stats = s->d_stats;
    return 0;
}
int nb200_sampler_geometry(nb200_sampler* s, int32_t* tpc, int32_t* block, int32_t* grid) {
    if (!s) return fail(NB200_EINVAL, "null sampler");
    if (tpc) *tpc = s->sub ? s->sub : 32 * s->W;
    if (block) *block = s->block;
    if (grid) *grid = s->grid;
    return 0;
}

int nb200_sampler_is_pipelined(nb200_sampler* s) { return s && s->piped ? 1 : 0; }

int nb200_sampler_smem(nb200_sampler* s, int32_t* smem_slots, int32_t* bytes_per_chain) {
    if (!s) return fail(NB200_EINVAL, "null sampler");
    if (smem_slots) *smem_slots = s->smem_slots;
    if (bytes_per_chain) *bytes_per_chain = (int32_t)s->smem_per_chain;
    return 0;
}

int nb200_sampler_destroy(nb200_sampler* s) {
    if (!s) return 0;
    cudaSetDevice(s->device);
    if (s->state == RunState::Running) {
        int one = 1;
        if (s->host && s->host->stop) {
            __atomic_store_n(s->host->stop, 1, __ATOMIC_RELEASE);
        } else if (s->d_stop && s->side) {
            cudaMemcpyAsync(s->d_stop, &one, sizeof(int), cudaMemcpyHostToDevice, s->side);
            cudaStreamSynchronize(s->side);
        }
        if (s->stream) cudaStreamSynchronize(s->stream);
    }
    if (s->host) s->host->shutdown();  // after the kernel: a waiting chain needs its answer
    void* dev[] = {s->d_pool, s->d_var, s->d_wf, s->d_sc, s->d_draws, s->d_stats, s->d_grads,
                   s->d_mm, s->d_div, s->d_q0, s->d_init_mean, s->d_stop, s->d_lr_stds, s->d_lr_vals,
                   s->d_lr_vecs, s->d_lr_coef, s->d_lr_win, s->d_lr_mat, s->d_lr_cols, s->d_eig};
    for (void* p : dev) {
        if (!p) continue;
        if (s->stream) cudaFreeAsync(p, s->stream);  // back to the pool
        else cudaFree(p);
    }
    for (void* p : s->model_allocs) {
        if (s->model_allocs_pooled && s->stream) cudaFreeAsync(p, s->stream);
        else cudaFree(p);
    }
    s->model_allocs.clear();
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->d_tape) cudaFree(s->d_tape);
    void* host[] = {s->h_draws, s->h_stats, s->h_grads, s->h_mm};
    for (void* p : host)
        if (p) cudaFreeHost(p);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->l2_persist) {
        // the set-aside is device-wide: hand the L2 back when the last sampler that uses it goes
        cudaCtxResetPersistingL2Cache();
        if (g_persist_users[s->device & 15].fetch_sub(1) == 1) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
        cudaGetLastError();
    }
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->side) cudaStreamDestroy(s->side);
    delete s;
    return 0;
}

}  // extern "C"
// ------------------------------------------------------ component entry points
template <class M>
static int component_run(const nb200_model_desc* model, int device, uint64_t n, int mode,
                         const double* q, const double* p, const double* g, const double* var,
                         const double* p_sum, const double* eps, const int32_t* dir,
                         const int64_t* idx, double* q_out, double* p_out, double* g_out,
                         double* p_sum_out, double* logp_out, double* kin_out, int32_t* rc_out) {
    CU(cudaSetDevice(device));
    int W = g_threads_per_chain.load() > 0 ? g_threads_per_chain.load() / 32 : 1;
    if (W < 1 || W > 32 || (W & (W - 1))) return fail(NB200_EINVAL, "bad threads per chain");
    const int D = (int)model->dim, Dp = (D + 3) / 4 * 4, NS = 2;
    KParams<M> P;
    std::memset(&P, 0, sizeof(P));
    nb200_settings_default(&P.st);
    P.st.max_energy_error = INFINITY;
    std::vector<void*> keep;
    int rc = build_model_data(*model, 32 * W, P.mdata, keep);
    if (rc != 0) return rc;
    P.D = D; P.Dp = Dp; P.NS = NS; P.n_chains = n;
    P.stage_loads = g_stage_loads.load();
    const size_t vecb = sizeof(double) * (size_t)Dp;
    double *d_pool, *d_var, *d_scal, *d_out;
    CU(cudaMalloc((void**)&d_pool, n * NS * 4 * vecb)); keep.push_back(d_pool);
    CU(cudaMalloc((void**)&d_var, n * vecb)); keep.push_back(d_var);
    CU(cudaMalloc((void**)&d_scal, n * 4 * sizeof(double))); keep.push_back(d_scal);
    CU(cudaMalloc((void**)&d_out, n * 4 * sizeof(double))); keep.push_back(d_out);
    CU(cudaMemset(d_pool, 0, n * NS * 4 * vecb));
    std::vector<double> hv(n * Dp, 1.0), hs(n * 4, 0.0);
    auto upload = [&](const double* src, int comp) -> int {
        if (!src) return 0;
        for (uint64_t c = 0; c < n; ++c)
            CU(cudaMemcpy(d_pool + ((c * NS + 0) * 4 + comp) * (size_t)Dp, src + c * D,
                          sizeof(double) * D, cudaMemcpyHostToDevice));
        return 0;
    };
    if ((rc = upload(q, VQ)) || (rc = upload(p, VP)) || (rc = upload(g, VG)) || (rc = upload(p_sum, VS)))
        return rc;
    if (var)
        for (uint64_t c = 0; c < n; ++c)
            for (int i = 0; i < D; ++i) hv[c * Dp + i] = var[c * D + i];
    CU(cudaMemcpy(d_var, hv.data(), n * vecb, cudaMemcpyHostToDevice));
    for (uint64_t c = 0; c < n; ++c) {
        hs[c * 4 + 0] = eps ? eps[c] : 0.0;
        hs[c * 4 + 1] = dir ? (double)dir[c] : 1.0;
        hs[c * 4 + 2] = idx ? (double)idx[c] : 0.0;
    }
    CU(cudaMemcpy(d_scal, hs.data(), n * 4 * sizeof(double), cudaMemcpyHostToDevice));
    P.pool = d_pool; P.var = d_var;
    const size_t smem = smem_for<M>(W, P.mdata, Dp);
    CU(launch_component<M>(W, P, mode, d_scal, d_out, smem, (unsigned)n));
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    std::vector<double> ho(n * 4);
    CU(cudaMemcpy(ho.data(), d_out, n * 4 * sizeof(double), cudaMemcpyDeviceToHost));
    auto download = [&](double* dst, int slot, int comp) -> int {
        if (!dst) return 0;
        for (uint64_t c = 0; c < n; ++c)
            CU(cudaMemcpy(dst + c * D, d_pool + ((c * NS + slot) * 4 + comp) * (size_t)Dp,
                          sizeof(double) * D, cudaMemcpyDeviceToHost));
        return 0;
    };
    const int oslot = mode == 0 ? 0 : 1;
    if ((rc = download(q_out, oslot, VQ)) || (rc = download(p_out, oslot, VP)) ||
        (rc = download(g_out, oslot, VG)) || (rc = download(p_sum_out, oslot, VS)))
        return rc;
    for (uint64_t c = 0; c < n; ++c) {
        if (logp_out) logp_out[c] = ho[c * 4 + 0];
        if (kin_out) kin_out[c] = ho[c * 4 + 1];
        if (rc_out) rc_out[c] = (int32_t)ho[c * 4 + 2];
    }
    for (void* ptr : keep) cudaFree(ptr);
    return 0;
}

extern "C" {
#define DISPATCH_MODEL(CALL)                                                   \
    switch (model->kind) {                                                     \
    case NB200_MODEL_NORMAL: return CALL(NormalModel);                         \
    case NB200_MODEL_FUNNEL: return CALL(FunnelModel);                         \
    case NB200_MODEL_RADON: return CALL(RadonModel);                           \
    case NB200_MODEL_CUSTOM: return CALL(CustomModel);                         \
    case NB200_MODEL_HOST:                                                     \
        return fail(NB200_EINVAL, "component entry points evaluate DEVICE densities; " \
                                  "a host plug-in is called directly");                \
    }                                                                          \
    return fail(NB200_EINVAL, "unknown model kind");

int nb200_custom_model_compile(const nb200_model_desc* model, int threads_per_chain,
                               int dims_per_thread, char* log, size_t log_len) {
    if (log && log_len) log[0] = 0;
    nb200_settings st;
    nb200_settings_default(&st);
    if (validate(&st, model) != 0) return NB200_EINVAL;
    if (model->kind != NB200_MODEL_CUSTOM) return fail(NB200_EINVAL, "not a custom model");
    const int W = threads_per_chain / 32;
    if (threads_per_chain % 32 || W < 1 || W > 32 || (W & (W - 1)))
        return fail(NB200_EINVAL, "threads per chain must be 32..1024, power of two");
    const int nit = dims_per_thread > 0 ? supported_nit<CustomModel>(W, dims_per_thread) : 0;
    if (nit != dims_per_thread)
        return fail(NB200_EINVAL, "no unrolled kernel for this (threads, dims per thread) pair");
    if (custom_compile(custom_register(model->cuda_source), W, nit) != 0) {
        if (log && log_len) std::snprintf(log, log_len, "%s", custom_last_log());
        return fail(NB200_ECOMPILE, custom_last_log());
    }
    return 0;
}

int nb200_logp_grad(const nb200_model_desc* model, int device, uint64_t n, const double* q,
                    double* logp, double* grad, int32_t* rc) {
    nb200_settings st;
    nb200_settings_default(&st);
    if (validate(&st, model) != 0) return NB200_EINVAL;
    if (nb200_device_count() < 1) return fail(NB200_ECUDA, "no CUDA device available");
    if (!q || n < 1) return fail(NB200_EINVAL, "null/empty input");
#define CALL(M)                                                                               \
    component_run<M>(model, device, n, 0, q, nullptr, nullptr, nullptr, nullptr, nullptr,     \
                     nullptr, nullptr, nullptr, nullptr, grad, nullptr, logp, nullptr, rc)
    DISPATCH_MODEL(CALL)
#undef CALL
}

int nb200_leapfrog(const nb200_model_desc* model, int device, uint64_t n, const double* q,
                   const double* p, const double* g, const double* var, const double* p_sum,
                   const double* eps, const int32_t* dir, const int64_t* idx, double* q_out,
                   double* p_out, double* g_out, double* p_sum_out, double* logp_out,
                   double* kinetic_out, int32_t* rc) {
    nb200_settings st;
    nb200_settings_default(&st);
    if (validate(&st, model) != 0) return NB200_EINVAL;
    if (nb200_device_count() < 1) return fail(NB200_ECUDA, "no CUDA device available");
    if (!q || !p || !g || !var || !p_sum || !eps || n < 1) return fail(NB200_EINVAL, "null/empty input");
#define CALL(M)                                                                                \
    component_run<M>(model, device, n, 1, q, p, g, var, p_sum, eps, dir, idx, q_out, p_out,    \
                     g_out, p_sum_out, logp_out, kinetic_out, rc)
    DISPATCH_MODEL(CALL)
#undef CALL
}

}  // extern "C"

// ---- low-rank metric at the component seam (tests): one warp refreshes the metric from a
// window of draws / gradients, then applies M^-1 and M^1/2 to the caller's vectors
__global__ void __launch_bounds__(32)
    lr_component_kernel(LrState L, int D, int Dp, double gamma, double cutoff, int n_vec,
                        const double* p_in, double* v_out, double* z_io, int* out) {
    GroupCuda<1> g;
    g.tid = threadIdx.x;
    g.red = nullptr;
    const bool ok = lr_update(g, L, D, Dp, gamma, cutoff);
    for (int v = 0; v < n_vec; ++v) {
        lr_velocity(g, L, D, Dp, p_in + (size_t)v * Dp, v_out + (size_t)v * Dp);
        lr_momentum(g, L, D, Dp, z_io + (size_t)v * Dp);
    }
    if (g.tid == 0) {
        out[0] = ok ? 1 : 0;
        out[1] = L.k;
    }
}

extern "C" int nb200_lowrank_component(int device, uint64_t dim, uint64_t n, const double* draws,
                                       const double* grads, double gamma, double cutoff,
                                       uint64_t max_rank, uint64_t n_vec, const double* p,
                                       double* v_out, const double* z, double* momentum_out,
                                       double* stds_out, double* vals_out, double* vecs_out,
                                       uint64_t* rank_out) {
    if (!draws || !grads || dim < 1 || dim > 1024 || n < 1 || max_rank < 1)
        return fail(NB200_EINVAL, "bad low-rank component input");
    if (nb200_device_count() < 1) return fail(NB200_ECUDA, "no CUDA device available");
    CU(cudaSetDevice(device));
    const int D = (int)dim, Dp = (D + 3) / 4 * 4;
    const int R = (int)(max_rank < dim ? max_rank : dim);
    const size_t vecb = sizeof(double) * (size_t)Dp;
    std::vector<void*> keep;
    auto dalloc = [&](size_t bytes) -> double* {
        void* ptr = nullptr;
        if (cudaMalloc(&ptr, bytes ? bytes : 8) != cudaSuccess) return nullptr;
        cudaMemset(ptr, 0, bytes ? bytes : 8);
        keep.push_back(ptr);
        return (double*)ptr;
    };
    auto cleanup = [&]() {
        for (void* ptr : keep) cudaFree(ptr);
    };
    LrState L;
    std::memset(&L, 0, sizeof(L));
    L.stds = dalloc(vecb); L.vals = dalloc(sizeof(double) * R); L.vecs = dalloc(vecb * R);
    L.coef = dalloc(sizeof(double) * R); L.win = dalloc(vecb * 2 * n);
    L.matL = dalloc(vecb * D); L.matW = dalloc(vecb * D); L.cols = dalloc(vecb * 6);
    double* d_p = dalloc(vecb * (n_vec ? n_vec : 1));
    double* d_v = dalloc(vecb * (n_vec ? n_vec : 1));
    double* d_z = dalloc(vecb * (n_vec ? n_vec : 1));
    int* d_out = (int*)dalloc(2 * sizeof(int));
    if (!d_out || !L.win || !L.matW) {
        cleanup();
        return fail(NB200_ECUDA, "cudaMalloc failed");
    }
    L.cap = (int)n; L.len = (int)n; L.split = 0; L.head = 0; L.k = 0; L.max_rank = R;
    std::vector<double> h((size_t)n * 2 * Dp, 0.0), ones(Dp, 1.0);
    for (uint64_t j = 0; j < n; ++j)
        for (int i = 0; i < D; ++i) {
            h[(j * 2 + 0) * Dp + i] = draws[j * dim + i];
            h[(j * 2 + 1) * Dp + i] = grads[j * dim + i];
        }
    cudaMemcpy(L.win, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(L.stds, ones.data(), vecb, cudaMemcpyHostToDevice);
    for (uint64_t v = 0; v < n_vec; ++v) {
        if (p) cudaMemcpy(d_p + v * Dp, p + v * dim, sizeof(double) * D, cudaMemcpyHostToDevice);
        if (z) cudaMemcpy(d_z + v * Dp, z + v * dim, sizeof(double) * D, cudaMemcpyHostToDevice);
    }
    lr_component_kernel<<<1, 32>>>(L, D, Dp, gamma, cutoff, (int)n_vec, d_p, d_v, d_z, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cleanup();
        return fail(NB200_ECUDA, std::string("lr_component_kernel: ") + cudaGetErrorString(e));
    }
    int out[2] = {0, 0};
    cudaMemcpy(out, d_out, sizeof(out), cudaMemcpyDeviceToHost);
    for (uint64_t v = 0; v < n_vec; ++v) {
        if (v_out) cudaMemcpy(v_out + v * dim, d_v + v * Dp, sizeof(double) * D, cudaMemcpyDeviceToHost);
        if (momentum_out) cudaMemcpy(momentum_out + v * dim, d_z + v * Dp, sizeof(double) * D, cudaMemcpyDeviceToHost);
    }
    if (stds_out) cudaMemcpy(stds_out, L.stds, sizeof(double) * D, cudaMemcpyDeviceToHost);
    if (vals_out) cudaMemcpy(vals_out, L.vals, sizeof(double) * out[1], cudaMemcpyDeviceToHost);
    if (vecs_out)
        for (int k = 0; k < out[1]; ++k)
            cudaMemcpy(vecs_out + (size_t)k * dim, L.vecs + (size_t)k * Dp, sizeof(double) * D, cudaMemcpyDeviceToHost);
    if (rank_out) *rank_out = (uint64_t)out[1];
    cleanup();
    return out[0] ? 0 : fail(NB200_EINVAL, "low-rank refresh broke down (window too short or not positive definite)");
}
