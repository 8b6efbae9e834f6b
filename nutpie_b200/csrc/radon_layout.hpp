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

namespace nb200 {

struct RadonLayout {
    int J = 0, N = 0, T = 0, n_steps = 0, G = 0;
    std::vector<int32_t> meta;         // [n_steps][T]  (2*county+floor) | end-of-group flag
    std::vector<double> y;             // [n_steps][T]
    std::vector<int32_t> group_base;   // [T]
    std::vector<int32_t> group_start;  // [2J+1]
};

inline RadonLayout build_radon_layout(int n_obs, int n_county, const double* y,
                                      const int32_t* county, const uint8_t* floor, int T) {
    constexpr int32_t kEndFlag = 1 << 30;
    RadonLayout L;
    L.J = n_county;
    L.N = n_obs;
    L.T = T;
    std::vector<int> order(n_obs);
    std::iota(order.begin(), order.end(), 0);
    auto key = [&](int o) { return 2 * county[o] + (floor[o] ? 1 : 0); };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) < key(b); });
    const int per = (n_obs + T - 1) / T;  // observations per thread
    L.n_steps = ((per > 0 ? per : 1) + 3) / 4 * 4;  // padded: the device loop is unrolled by 4
    L.meta.assign((size_t)L.n_steps * T, -1);
    L.y.assign((size_t)L.n_steps * T, 0.0);
    L.group_base.assign(T, 0);
    L.group_start.assign(2 * n_county + 1, 0);
    std::vector<int> group_key;
    for (int t = 0; t < T; ++t) {
        L.group_base[t] = (int32_t)group_key.size();
        for (int j = 0; j < per; ++j) {
            const int pos = t * per + j;
            if (pos >= n_obs) break;
            const int o = order[pos];
            const int kcur = key(o);
            const bool last_of_range = (j + 1 == per) || (pos + 1 >= n_obs);
            const bool ends = last_of_range || key(order[pos + 1]) != kcur;
            L.meta[(size_t)j * T + t] = kcur | (ends ? kEndFlag : 0);
            L.y[(size_t)j * T + t] = y[o];
            if (ends) group_key.push_back(kcur);
        }
    }
    L.G = (int)group_key.size();
    // groups are ordered by key because the observations are; pair k owns the
    // contiguous slot range [group_start[k], group_start[k+1])
    size_t r = 0;
    for (int k = 0; k <= 2 * n_county; ++k) {
        while (r < group_key.size() && group_key[r] < k) ++r;
        L.group_start[k] = (int32_t)r;
    }
    if (L.G == 0) L.G = 1;
    return L;
}

}  // namespace nb200
