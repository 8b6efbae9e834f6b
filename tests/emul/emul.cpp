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

template <class M, int NIT>
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
    std::vector<double> pool((size_t)P.NS * 4 * P.Dp), var(P.Dp), wf(8 * (size_t)P.Dp);
    std::vector<ChainScalars> sc(n_chains);
    std::memset(sc.data(), 0, sizeof(ChainScalars) * n_chains);
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
            ChainCtx<M, GroupSerial, NIT> ctx;
            std::memset((void*)&ctx, 0, sizeof(ctx));
            ctx.P = &P; ctx.md = P.mdata; ctx.sh = &sh; ctx.msm = msm.data();
            ctx.front = stage.data(); ctx.front_slot = -1; ctx.sweep_rev = false; ctx.n_parked = 0; ctx.defer_acc = false;
            std::fill(stage.begin(), stage.end(), -555.0);
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
    switch (dim) {
    case 1: return run_all_nit<M, 1>(st, md, dim, rest...);
    case 2: return run_all_nit<M, 2>(st, md, dim, rest...);
    case 5: return run_all_nit<M, 6>(st, md, dim, rest...);  // NIT * T > D: predicated tail
    case 6: return run_all_nit<M, 6>(st, md, dim, rest...);
    default: return run_all_nit<M, 0>(st, md, dim, rest...);
    }
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
