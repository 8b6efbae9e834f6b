// radon_layout.hpp — host-side "compile" step of the radon density: turns the
// user's observation arrays (any order; README.md:45-51 factorises counties in
// file order) into the per-thread run layout RadonModel reads (models.cuh).
// This is the device analogue of what compile_pymc_model does when it bakes
// `pm.Data`/observed arrays into the numba cfunc's user_data
// (python/nutpie/compile_pymc.py:248-269, 307-319).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "models.cuh"

namespace nb200 {

struct RadonLayout {
    int J = 0, N = 0, T = 0, n_steps = 0, G = 0, kmax = 1;
    std::vector<RadonObs> obs;           // [n_steps][T]
    std::vector<int32_t> group_base;     // [T]
    // [J][4]: for the county's (floor 0, floor 1) pairs, where the pair's LAST piece ends and
    // where the previous non-empty pair's last piece ends — {slot | prev_slot << 16,
    // thread | prev_thread << 16} each; an empty pair has slot = prev_slot (sum 0)
    std::vector<uint32_t> group_list;
};

inline RadonLayout build_radon_layout(int n_obs, int n_county, const double* y,
                                      const int32_t* county, const uint8_t* floor, int T) {
    RadonLayout L;
    L.J = n_county;
    L.N = n_obs;
    L.T = T;
    std::vector<int> order(n_obs);
    std::iota(order.begin(), order.end(), 0);
    auto key = [&](int o) { return 2 * county[o] + (floor[o] ? 1 : 0); };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) < key(b); });
    const int per = (n_obs + T - 1) / T;  // observations per thread
    L.n_steps = per > 0 ? per : 1;  // the device loop takes four at a time, then the remainder
    const int32_t dummy_mu = (int32_t)(2 * n_county) * 8;  // byte offset of mu[2J] (always 0)
    L.obs.assign((size_t)L.n_steps * T, RadonObs{0.0, dummy_mu, 0});
    L.group_base.assign(T, 0);
    std::vector<std::vector<std::pair<int, int>>> pieces(2 * n_county);  // (slot, prev or -1)
    int slot = 0;
    for (int t = 0; t < T; ++t) {
        L.group_base[t] = slot;
        int prev = -1;
        for (int j = 0; j < per; ++j) {
            const int pos = t * per + j;
            if (pos >= n_obs) break;
            const int o = order[pos];
            const int kcur = key(o);
            const bool last_of_range = (j + 1 == per) || (pos + 1 >= n_obs);
            const bool ends = last_of_range || key(order[pos + 1]) != kcur;
            L.obs[(size_t)j * T + t] = RadonObs{y[o], (int32_t)(kcur * 8) | (ends ? 1 : 0), 0};
            if (ends) {
                pieces[kcur].push_back({slot, prev});
                prev = slot++;
            }
        }
    }
    L.G = slot;  // slot G itself (one past) always holds 0: "before the first observation"
    L.kmax = 1;
    for (auto& p : pieces) L.kmax = std::max<int>(L.kmax, (int)p.size());
    // A pair's sum of residuals = global prefix at its end - global prefix at the end of the
    // previous non-empty pair, where the global prefix at a group end = the owning thread's
    // running prefix (stored in the group's slot) + the sum of all earlier threads' ranges
    // (exclusive scan over the threads, slot T of that table = 0).
    std::vector<int> owner(slot + 1, T);  // thread that owns each group slot (slot G -> T)
    for (int t = 0; t < T; ++t) {
        const int hi = t + 1 < T ? L.group_base[t + 1] : slot;
        for (int g2 = L.group_base[t]; g2 < hi; ++g2) owner[g2] = t;
    }
    L.group_list.assign((size_t)4 * n_county, 0u);
    int prev_slot = L.G, prev_thread = T;
    for (int k = 0; k < 2 * n_county; ++k) {
        uint32_t a, b;
        if (pieces[k].empty()) {  // no observations: both ends coincide
            a = (uint32_t)L.G | ((uint32_t)L.G << 16);
            b = (uint32_t)T | ((uint32_t)T << 16);
        } else {
            const int last = pieces[k].back().first;
            a = (uint32_t)last | ((uint32_t)prev_slot << 16);
            b = (uint32_t)owner[last] | ((uint32_t)prev_thread << 16);
            prev_slot = last;
            prev_thread = owner[last];
        }
        L.group_list[(size_t)2 * k] = a;
        L.group_list[(size_t)2 * k + 1] = b;
    }
    return L;
}

}  // namespace nb200
