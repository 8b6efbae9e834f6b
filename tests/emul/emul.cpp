// tests/emul/emul.cpp — TEST INFRASTRUCTURE, never shipped or loaded by the product.
//
// Compiles the sampler core (nutpie_b200/csrc/nuts_core.cuh) as plain host C++
// with a one-thread group, so the iterative tree / slot bookkeeping / adaptation
// logic of the CUDA engine can be checked against the recursive oracle on a
// machine without a GPU.  The CUDA build uses the very same header with
// GroupCuda<W>; only the thread-group policy differs.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../nutpie_b200/csrc/nuts_core.cuh"
#include "../../nutpie_b200/csrc/radon_layout.hpp"

using namespace nb200;

template <class M, int NIT, bool LR = false>
static int run_all_nit(const nb200_settings* st, const typename M::Data& md, uint64_t dim,
                   uint64_t n_chains, uint64_t chain_id_offset, const double* q0,
                   const double* init_mean, const double* z_tape, double* draws, double* stats,
                   double* grads, double* mminv, uint64_t* total_steps, int max_per_launch, int smem_slots) {
    KParams<M> P;
    std::memset(&P, 0, sizeof(P));
    P.st = *st;
    P.mdata = md;
    P.D = (int)dim;
    P.Dp = (int)((dim + 3) / 4 * 4);
    P.NS = 3 * ((int)st->maxdepth + 1) + 3;
    if (P.NS > kMaxSlots) return NB200_EINVAL;
    P.n_chains = n_chains;
    P.chain_id_offset = chain_id_offset;
    P.n_total = st->num_tune + st->num_draws;
    P.n_rows = st->save_warmup ? P.n_total : st->num_draws;
    P.sdim = (st->store_dims && st->store_dims < dim) ? st->store_dims : dim;
    P.gdim = P.sdim;
    P.expand = 0;
    P.max_draws_per_launch = max_per_launch;
    P.smem_slots = smem_slots < P.NS ? smem_slots : P.NS;
    P.var_in_smem = smem_slots > 0;
    std::vector<double> pool((size_t)P.NS * (LR ? 5 : 4) * P.Dp), var(P.Dp), wf(8 * (size_t)P.Dp);
    std::vector<ChainScalars> sc(n_chains);
    std::memset(sc.data(), 0, sizeof(ChainScalars) * n_chains);
    // low-rank adaptation: one chain's metric / window / scratch (chains run one after another)
    const uint64_t lr_f = st->mass_matrix_switch_freq > st->early_mass_matrix_switch_freq
                              ? st->mass_matrix_switch_freq : st->early_mass_matrix_switch_freq;
    const int lr_cap = LR ? (int)(3 * lr_f + 2) : 1;
    const int lr_rank = LR ? (int)(st->mass_matrix_max_rank < dim ? st->mass_matrix_max_rank : dim) : 1;
    std::vector<double> lr_stds(P.Dp), lr_vals(lr_rank), lr_vecs((size_t)lr_rank * P.Dp), lr_coef(lr_rank);
    std::vector<double> lr_win(LR ? (size_t)lr_cap * 2 * P.Dp : 1), lr_mat(LR ? 2 * (size_t)P.D * P.Dp : 1);
    std::vector<double> lr_cols(6 * (size_t)P.Dp);
    P.sc = sc.data();
    P.draws = draws; P.stats = stats; P.grads = grads; P.mminv = mminv;
    P.q0 = q0; P.init_mean = init_mean; P.z_tape = z_tape;
    P.stop_flag = nullptr;
    std::vector<double> msm(M::smem_doubles(md, 1) + 1), stage(4 * (size_t)P.Dp + 1);
    std::vector<double> spool((size_t)P.smem_slots * 4 * P.Dp + 1), svar(P.Dp);
    ChainShared sh;
    uint64_t steps = 0;
    int err = 0;
    for (uint64_t c = 0; c < n_chains; ++c) {
        // the pool is per chain; chains run one after another here, but a chain may
        // be resumed over several "launches" (max_per_launch) to exercise pause/resume
        std::fill(pool.begin(), pool.end(), 0.0);
        for (;;) {
            // the shared-memory tier does not survive a launch: scramble it
            std::fill(spool.begin(), spool.end(), -777.0);
            std::fill(svar.begin(), svar.end(), -777.0);
            ChainCtx<M, GroupSerial, NIT, LR> ctx;
            std::memset((void*)&ctx, 0, sizeof(ctx));
            ctx.P = &P; ctx.md = P.mdata; ctx.sh = &sh; ctx.msm = msm.data();
            ctx.front = stage.data(); ctx.front_slot = -1; ctx.sweep_rev = false; ctx.n_parked = 0; ctx.defer_acc = false;
            std::fill(stage.begin(), stage.end(), -555.0);
            if (LR) {
                ctx.lr.stds = lr_stds.data(); ctx.lr.vals = lr_vals.data(); ctx.lr.vecs = lr_vecs.data();
                ctx.lr.coef = lr_coef.data(); ctx.lr.win = lr_win.data(); ctx.lr.matL = lr_mat.data();
                ctx.lr.matW = lr_mat.data() + (size_t)P.D * P.Dp; ctx.lr.cols = lr_cols.data();
                ctx.lr.cap = lr_cap; ctx.lr.max_rank = lr_rank;
            }
            ctx.D = P.D; ctx.Dp = P.Dp; ctx.NS = P.NS;
            ctx.chain_local = c;
            ctx.chain_gid = (uint32_t)(chain_id_offset + c);
            ctx.pool = pool.data(); ctx.varg = var.data(); ctx.wf = wf.data();
            ctx.var = P.var_in_smem ? svar.data() : var.data();
            ctx.spool = spool.data(); ctx.smem_slots = P.smem_slots;
            ctx.run();
            if (sc[c].status == 2 || sc[c].status < 0) break;
        }
        if (sc[c].status < 0) err = sc[c].status;
        steps += sc[c].total_steps;
    }
    if (total_steps) *total_steps = steps;
    return err;
}

// the unrolled (NIT > 0) code paths are exercised when one thread's trip count is small
template <class M, class... A>
static int run_all(const nb200_settings* st, const typename M::Data& md, uint64_t dim, A... rest) {
    if (st->adaptation == 1) return run_all_nit<M, 0, true>(st, md, dim, rest...);
    switch (dim) {
    case 1: return run_all_nit<M, 1>(st, md, dim, rest...);
    case 2: return run_all_nit<M, 2>(st, md, dim, rest...);
    case 5: return run_all_nit<M, 6>(st, md, dim, rest...);  // NIT * T > D: predicated tail
    case 6: return run_all_nit<M, 6>(st, md, dim, rest...);
    default: return run_all_nit<M, 0>(st, md, dim, rest...);
    }
}

// ---------------------------------------------------------------------------------------------
// Multi-lane emulation: T logical threads of ONE chain run the very code the GPU runs for a
// T-thread group (per-thread ChainCtx, shared pool / front / scratch), as cooperative fibers that
// hand over at every sync() / reduce() in lane order.  Between two hand-overs a lane runs alone,
// which is one of the schedules independent thread scheduling allows: a missing barrier in the
// core shows up as a wrong result here.  reduce() adds in the association order of
// GroupCuda::warp_reduce (xor butterfly) followed by the fixed-order cross-warp sum.
#include <ucontext.h>

struct LaneSched {
    int T = 0, cur = 0;
    std::vector<ucontext_t> ctx;
    ucontext_t main_ctx;
    std::vector<std::vector<char>> stacks;
    std::vector<char> alive;
    std::vector<double> red;  // [T][kMaxRed]
    static constexpr int kMaxRed = 16;
    void (*body)(int lane, void* arg) = nullptr;
    void* arg = nullptr;
    // Order in which the lanes run between two barriers: 0 ascending, 1 descending, >= 2 a new
    // pseudo-random permutation after every barrier (seeded by the value).  Every order is a legal
    // schedule, so the results must not depend on it.
    int order = 0;
    std::vector<int> perm;
    int pos = 0;
    uint64_t lcg = 0;
    void new_epoch() {
        if (order < 2) return;
        for (int i = T - 1; i > 0; --i) {
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            std::swap(perm[i], perm[(int)((lcg >> 33) % (uint64_t)(i + 1))]);
        }
    }
    void yield() {
        int n_alive = 0;
        for (int l = 0; l < T; ++l) n_alive += alive[l];
        if (n_alive <= 1) return;
        do {
            if (++pos >= T) {
                pos = 0;
                new_epoch();
            }
        } while (!alive[perm[pos]]);
        const int to = perm[pos];
        if (to == cur) return;
        const int from = cur;
        cur = to;
        swapcontext(&ctx[from], &ctx[to]);
    }
};
static LaneSched* g_sched = nullptr;
static void lane_trampoline(int lane) {
    g_sched->body(lane, g_sched->arg);
    g_sched->alive[lane] = 0;  // back to run_lanes() through uc_link
}
static void run_lanes(LaneSched& S, int T, void (*body)(int, void*), void* arg) {
    S.T = T; S.body = body; S.arg = arg;
    S.ctx.assign(T, ucontext_t());
    S.stacks.assign(T, std::vector<char>(256 * 1024));
    S.alive.assign(T, 1);
    S.red.assign((size_t)T * LaneSched::kMaxRed, 0.0);
    S.perm.resize(T);
    for (int l = 0; l < T; ++l) S.perm[l] = S.order == 1 ? T - 1 - l : l;
    S.lcg = 0x9E3779B97F4A7C15ull * (uint64_t)(S.order + 1);
    S.new_epoch();
    S.pos = 0;
    g_sched = &S;
    for (int l = 0; l < T; ++l) {
        getcontext(&S.ctx[l]);
        S.ctx[l].uc_stack.ss_sp = S.stacks[l].data();
        S.ctx[l].uc_stack.ss_size = S.stacks[l].size();
        S.ctx[l].uc_link = &S.main_ctx;
        makecontext(&S.ctx[l], (void (*)())lane_trampoline, 1, l);
    }
    for (;;) {
        int next = -1;
        for (int i = 0; i < T; ++i)
            if (S.alive[S.perm[i]]) { next = S.perm[i]; S.pos = i; break; }
        if (next < 0) break;
        S.cur = next;
        swapcontext(&S.main_ctx, &S.ctx[next]);
    }
    g_sched = nullptr;
}

template <int T>
struct GroupLanes {
    static constexpr int kThreads = T;
    int tid;
    LaneSched* s;
    int size() const { return T; }
    void sync() const { s->yield(); }
    template <int N>
    void reduce(double (&v)[N]) const {
        static_assert(N <= LaneSched::kMaxRed, "reduction too wide for the emulation");
        double* red = s->red.data();
        for (int i = 0; i < N; ++i) red[(size_t)tid * LaneSched::kMaxRed + i] = v[i];
        s->yield();  // every lane has published its partials
        for (int i = 0; i < N; ++i) {
            double total = 0.0;
            for (int w = 0; w < T / 32; ++w) {  // fixed-order sum over the warps of the group
                double x[32], y[32];
                for (int l = 0; l < 32; ++l) x[l] = red[(size_t)(w * 32 + l) * LaneSched::kMaxRed + i];
                for (int lv = 0; lv < 5; ++lv) {
                    // N < 4: offsets 16, 8, 4, 2, 1; N >= 4 (recursive halving): 1, 2, 4, 8, 16
                    const int off = N < 4 ? (16 >> lv) : (1 << lv);
                    for (int l = 0; l < 32; ++l) y[l] = x[l] + x[l ^ off];
                    for (int l = 0; l < 32; ++l) x[l] = y[l];
                }
                total = (T == 32) ? x[0] : total + x[0];
            }
            v[i] = total;
        }
        s->yield();  // every lane has read: the scratch may be reused
    }
    // GroupCuda::exclusive_scan in the same association order: Hillis-Steele inside each warp,
    // then the totals of the earlier warps added in warp order
    double exclusive_scan(double x) const {
        double* red = s->red.data();
        red[(size_t)tid * LaneSched::kMaxRed] = x;
        s->yield();
        const int w = tid / 32, lane = tid % 32;
        auto warp_incl = [&](int ww, double* out) {
            double v[32], y[32];
            for (int l = 0; l < 32; ++l) v[l] = red[(size_t)(ww * 32 + l) * LaneSched::kMaxRed];
            for (int off = 1; off < 32; off <<= 1) {
                for (int l = 0; l < 32; ++l) y[l] = l >= off ? v[l] + v[l - off] : v[l];
                for (int l = 0; l < 32; ++l) v[l] = y[l];
            }
            for (int l = 0; l < 32; ++l) out[l] = v[l];
        };
        double mine[32];
        warp_incl(w, mine);
        double excl = lane == 0 ? 0.0 : mine[lane - 1];
        if (T > 32) {
            double base = 0.0;
            for (int ww = 0; ww < w; ++ww) {
                double o[32];
                warp_incl(ww, o);
                base += o[31];
            }
            excl = base + excl;
        }
        s->yield();
        return excl;
    }
};

template <class M, int T, int NIT, bool LR = false>
struct LaneLaunch {
    const KParams<M>* P;
    ChainShared* sh;
    double *msm, *front, *pool, *var, *wf, *spool, *svar;
    uint64_t chain, chain_id_offset;
    LaneSched* sched;
    LrState lr0;  // low-rank adaptation: the chain's buffers (every lane keeps its own counters)
    static void body(int lane, void* arg) {
        LaneLaunch& L = *static_cast<LaneLaunch*>(arg);
        const KParams<M>& P = *L.P;
        ChainCtx<M, GroupLanes<T>, NIT, LR> ctx;
        std::memset((void*)&ctx, 0, sizeof(ctx));
        ctx.lr = L.lr0;
        ctx.g.tid = lane;
        ctx.g.s = L.sched;
        ctx.P = &P; ctx.md = P.mdata; ctx.sh = L.sh; ctx.msm = L.msm;
        ctx.front = L.front; ctx.front_slot = -1; ctx.sweep_rev = false; ctx.n_parked = 0; ctx.defer_acc = false;
        ctx.D = P.D; ctx.Dp = P.Dp; ctx.NS = P.NS;
        ctx.chain_local = L.chain;
        ctx.chain_gid = (uint32_t)(L.chain_id_offset + L.chain);
        ctx.pool = L.pool; ctx.varg = L.var; ctx.wf = L.wf;
        ctx.var = P.var_in_smem ? L.svar : L.var;
        ctx.spool = L.spool; ctx.smem_slots = P.smem_slots;
        ctx.mL = ctx.mR = ctx.mD = ctx.tL = ctx.tR = ctx.tD = -1;
        ctx.run();
    }
};

template <class M, int T, int NIT, bool LR = false>
static int run_lanes_nit(const nb200_settings* st, const typename M::Data& md, uint64_t dim,
                         uint64_t n_chains, uint64_t chain_id_offset, double* draws, double* stats,
                         uint64_t* total_steps, int max_per_launch, int smem_slots, int lane_order,
                         int expand) {
    KParams<M> P;
    std::memset(&P, 0, sizeof(P));
    P.st = *st;
    P.mdata = md;
    P.D = (int)dim;
    P.Dp = (int)((dim + 3) / 4 * 4);
    P.NS = 3 * ((int)st->maxdepth + 1) + 3;
    if (P.NS > kMaxSlots) return NB200_EINVAL;
    P.n_chains = n_chains;
    P.chain_id_offset = chain_id_offset;
    P.n_total = st->num_tune + st->num_draws;
    P.n_rows = st->save_warmup ? P.n_total : st->num_draws;
    P.expand = expand;  // draws hold the expanded vector (Model::expand run by the lanes)
    P.sdim = expand ? (uint64_t)M::expanded_dim((int)dim) : dim;
    P.gdim = dim;
    P.max_draws_per_launch = max_per_launch;
    P.smem_slots = smem_slots < P.NS ? smem_slots : P.NS;
    P.var_in_smem = smem_slots > 0;
    std::vector<double> pool((size_t)P.NS * (LR ? 5 : 4) * P.Dp), var(P.Dp), wf(8 * (size_t)P.Dp);
    std::vector<ChainScalars> sc(n_chains);
    std::memset(sc.data(), 0, sizeof(ChainScalars) * n_chains);
    P.sc = sc.data();
    P.draws = draws; P.stats = stats;
    const uint64_t lr_f = st->mass_matrix_switch_freq > st->early_mass_matrix_switch_freq
                              ? st->mass_matrix_switch_freq : st->early_mass_matrix_switch_freq;
    const int lr_cap = LR ? (int)(3 * lr_f + 2) : 1;
    const int lr_rank = LR ? (int)(st->mass_matrix_max_rank < dim ? st->mass_matrix_max_rank : dim) : 1;
    std::vector<double> lr_stds(P.Dp), lr_vals(lr_rank), lr_vecs((size_t)lr_rank * P.Dp), lr_coef(lr_rank);
    std::vector<double> lr_win(LR ? (size_t)lr_cap * 2 * P.Dp : 1), lr_mat(LR ? 2 * (size_t)P.D * P.Dp : 1);
    std::vector<double> lr_cols(6 * (size_t)P.Dp);
    LrState lr0;
    std::memset(&lr0, 0, sizeof(lr0));
    lr0.stds = lr_stds.data(); lr0.vals = lr_vals.data(); lr0.vecs = lr_vecs.data(); lr0.coef = lr_coef.data();
    lr0.win = lr_win.data(); lr0.matL = lr_mat.data(); lr0.matW = lr_mat.data() + (LR ? (size_t)P.D * P.Dp : 0);
    lr0.cols = lr_cols.data(); lr0.cap = lr_cap; lr0.max_rank = lr_rank;
    std::vector<double> msm(M::smem_doubles(md, T) + 1), front(4 * (size_t)P.Dp + 1);
    std::vector<double> spool((size_t)P.smem_slots * 4 * P.Dp + 1), svar(P.Dp);
    ChainShared sh;
    uint64_t steps = 0;
    int err = 0;
    for (uint64_t c = 0; c < n_chains; ++c) {
        std::fill(pool.begin(), pool.end(), 0.0);
        for (;;) {
            std::fill(spool.begin(), spool.end(), -777.0);
            std::fill(svar.begin(), svar.end(), -777.0);
            std::fill(front.begin(), front.end(), -555.0);
            LaneSched sched;
            sched.order = lane_order;
            LaneLaunch<M, T, NIT, LR> L{&P, &sh, msm.data(), front.data(), pool.data(), var.data(), wf.data(),
                                        spool.data(), svar.data(), c, chain_id_offset, &sched, lr0};
            run_lanes(sched, T, &LaneLaunch<M, T, NIT, LR>::body, &L);
            if (sc[c].status == 2 || sc[c].status < 0) break;
        }
        if (sc[c].status < 0) err = sc[c].status;
        steps += sc[c].total_steps;
    }
    if (total_steps) *total_steps = steps;
    return err;
}

// T lanes per chain with the trip count the product picks for (T, dim): ceil(dim / T), unrolled
extern "C" int emul_sample_lanes(const nb200_settings* st, const nb200_model_desc* model, int T,
                                 uint64_t n_chains, uint64_t chain_id_offset, double* draws,
                                 double* stats, uint64_t* total_steps, int max_per_launch,
                                 int smem_slots, int lane_order, int expand) {
    const uint64_t dim = model->dim;
    const int nit = (int)((dim + T - 1) / T);
#define LANES_CASE(MODEL, DATA, TT, NN)                                                          \
    if (T == TT && nit == NN)                                                                    \
        return run_lanes_nit<MODEL, TT, NN>(st, DATA, dim, n_chains, chain_id_offset, draws, stats, \
                                            total_steps, max_per_launch, smem_slots, lane_order, expand);
    if (st->adaptation == 1) {  // low-rank engine: a warp of lanes, run-time loops
        if (T != 32) return NB200_EINVAL;
        switch (model->kind) {
        case NB200_MODEL_NORMAL: {
            NormalModel::Data d{model->mu, 1.0 / (model->sigma * model->sigma)};
            return run_lanes_nit<NormalModel, 32, 0, true>(st, d, dim, n_chains, chain_id_offset, draws, stats,
                                                           total_steps, max_per_launch, 0, lane_order, expand);
        }
        case NB200_MODEL_FUNNEL: {
            FunnelModel::Data d{0};
            return run_lanes_nit<FunnelModel, 32, 0, true>(st, d, dim, n_chains, chain_id_offset, draws, stats,
                                                           total_steps, max_per_launch, 0, lane_order, expand);
        }
        case NB200_MODEL_RADON: {
            RadonLayout L = build_radon_layout(model->n_obs, model->n_county, model->y, model->county,
                                               model->floor, 32);
            RadonModel::Data d{L.J, L.N, L.n_steps, L.G, L.kmax, 32, 0, L.obs.data(), L.group_base.data(),
                               L.group_list.data()};
            return run_lanes_nit<RadonModel, 32, 0, true>(st, d, dim, n_chains, chain_id_offset, draws, stats,
                                                          total_steps, max_per_launch, 0, lane_order, expand);
        }
        }
        return NB200_EINVAL;
    }
    switch (model->kind) {
    case NB200_MODEL_NORMAL: {
        NormalModel::Data d{model->mu, 1.0 / (model->sigma * model->sigma)};
        LANES_CASE(NormalModel, d, 32, 1) LANES_CASE(NormalModel, d, 32, 2) LANES_CASE(NormalModel, d, 64, 1)
        if (T == 32) return run_lanes_nit<NormalModel, 32, 0>(st, d, dim, n_chains, chain_id_offset, draws, stats,
                                                              total_steps, max_per_launch, smem_slots, lane_order, expand);
        break;
    }
    case NB200_MODEL_FUNNEL: {
        FunnelModel::Data d{0};
        LANES_CASE(FunnelModel, d, 32, 1)
        break;
    }
    case NB200_MODEL_RADON: {
        RadonLayout L = build_radon_layout(model->n_obs, model->n_county, model->y, model->county,
                                           model->floor, T);
        RadonModel::Data d{L.J, L.N, L.n_steps, L.G, L.kmax, T, 0, L.obs.data(), L.group_base.data(),
                           L.group_list.data()};
        LANES_CASE(RadonModel, d, 32, 6) LANES_CASE(RadonModel, d, 64, 3) LANES_CASE(RadonModel, d, 128, 2)
        LANES_CASE(RadonModel, d, 32, 1)
        break;
    }
    }
#undef LANES_CASE
    return NB200_EINVAL;
}

extern "C" uint64_t emul_expanded_dim(const nb200_model_desc* model) {
    switch (model->kind) {
    case NB200_MODEL_NORMAL: return (uint64_t)NormalModel::expanded_dim((int)model->dim);
    case NB200_MODEL_FUNNEL: return (uint64_t)FunnelModel::expanded_dim((int)model->dim);
    case NB200_MODEL_RADON: return (uint64_t)RadonModel::expanded_dim((int)model->dim);
    }
    return 0;
}

extern "C" int emul_sample(const nb200_settings* st, const nb200_model_desc* model,
                           uint64_t n_chains, uint64_t chain_id_offset, const double* q0,
                           const double* init_mean, const double* z_tape, double* draws,
                           double* stats, double* grads, double* mminv, uint64_t* total_steps,
                           int max_per_launch, int smem_slots) {
    switch (model->kind) {
    case NB200_MODEL_NORMAL: {
        NormalModel::Data d{model->mu, 1.0 / (model->sigma * model->sigma)};
        return run_all<NormalModel>(st, d, model->dim, n_chains, chain_id_offset, q0, init_mean,
                                    z_tape, draws, stats, grads, mminv, total_steps, max_per_launch, smem_slots);
    }
    case NB200_MODEL_FUNNEL: {
        FunnelModel::Data d{0};
        return run_all<FunnelModel>(st, d, model->dim, n_chains, chain_id_offset, q0, init_mean,
                                    z_tape, draws, stats, grads, mminv, total_steps, max_per_launch, smem_slots);
    }
    case NB200_MODEL_RADON: {
        RadonLayout L = build_radon_layout(model->n_obs, model->n_county, model->y, model->county,
                                           model->floor, 1);
        RadonModel::Data d{L.J, L.N, L.n_steps, L.G, L.kmax, 1, 0, L.obs.data(), L.group_base.data(),
                           L.group_list.data()};
        return run_all<RadonModel>(st, d, model->dim, n_chains, chain_id_offset, q0, init_mean,
                                   z_tape, draws, stats, grads, mminv, total_steps, max_per_launch, smem_slots);
    }
    }
    return NB200_EINVAL;
}

// the low-rank refresh + metric application of lowrank.cuh on one host thread (the twin of
// nb200_lowrank_component): window [n][dim] -> stds, vals, vecs; v = M^-1 p, pm = M^1/2 z
extern "C" int emul_lowrank_component(uint64_t dim, uint64_t n, const double* draws, const double* grads,
                                      double gamma, double cutoff, uint64_t max_rank, uint64_t n_vec,
                                      const double* p, double* v_out, const double* z, double* pm_out,
                                      double* stds_out, double* vals_out, double* vecs_out,
                                      uint64_t* rank_out) {
    const int D = (int)dim, Dp = (D + 3) / 4 * 4;
    const int R = (int)(max_rank < dim ? max_rank : dim);
    std::vector<double> stds(Dp, 1.0), vals(R), vecs((size_t)R * Dp), coef(R), win((size_t)n * 2 * Dp),
        mat(2 * (size_t)D * Dp), cols(6 * (size_t)Dp), pv(Dp), vv(Dp);
    for (uint64_t j = 0; j < n; ++j)
        for (int i = 0; i < D; ++i) {
            win[(j * 2 + 0) * Dp + i] = draws[j * dim + i];
            win[(j * 2 + 1) * Dp + i] = grads[j * dim + i];
        }
    LrState L;
    L.stds = stds.data(); L.vals = vals.data(); L.vecs = vecs.data(); L.coef = coef.data();
    L.win = win.data(); L.matL = mat.data(); L.matW = mat.data() + (size_t)D * Dp; L.cols = cols.data();
    L.k = 0; L.len = (int)n; L.split = 0; L.head = 0; L.cap = (int)n; L.max_rank = R;
    GroupSerial g;
    g.tid = 0;
    const bool ok = lr_update(g, L, D, Dp, gamma, cutoff);
    for (uint64_t v = 0; v < n_vec; ++v) {
        if (p && v_out) {
            for (int i = 0; i < D; ++i) pv[i] = p[v * dim + i];
            lr_velocity(g, L, D, Dp, pv.data(), vv.data());
            for (int i = 0; i < D; ++i) v_out[v * dim + i] = vv[i];
        }
        if (z && pm_out) {
            for (int i = 0; i < D; ++i) pv[i] = z[v * dim + i];
            lr_momentum(g, L, D, Dp, pv.data());
            for (int i = 0; i < D; ++i) pm_out[v * dim + i] = pv[i];
        }
    }
    if (stds_out) for (int i = 0; i < D; ++i) stds_out[i] = stds[i];
    if (vals_out) for (int k = 0; k < L.k; ++k) vals_out[k] = vals[k];
    if (vecs_out)
        for (int k = 0; k < L.k; ++k)
            for (int i = 0; i < D; ++i) vecs_out[(size_t)k * dim + i] = vecs[(size_t)k * Dp + i];
    if (rank_out) *rank_out = (uint64_t)L.k;
    return ok ? 0 : 1;
}

// the tournament schedule of the one-sided Jacobi (lowrank.cuh): pairs of round t, for tests
extern "C" void emul_round_robin_pairs(int m, int t, int* a_out, int* b_out) {
    for (int k = 0; k < m / 2; ++k) lr_round_robin_pair(t, k, m, a_out[k], b_out[k]);
}
