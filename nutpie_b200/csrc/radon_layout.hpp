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
    int J = 0, N = 0, T = 0, n_steps = 0, R = 0;
    std::vector<int32_t> packed;     // [n_steps][T]
    std::vector<double> y;           // [n_steps][T]
    std::vector<int32_t> run_base;   // [T]
    std::vector<int32_t> run_start;  // [J+1]
};

inline RadonLayout build_radon_layout(int n_obs, int n_county, const double* y,
                                      const int32_t* county, const uint8_t* floor, int T) {
    RadonLayout L;
    L.J = n_county;
    L.N = n_obs;
    L.T = T;
    std::vector<int> order(n_obs);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](int a, int b) { return county[a] < county[b]; });
    const int per = (n_obs + T - 1) / T;  // observations per thread
    L.n_steps = ((per > 0 ? per : 1) + 3) / 4 * 4;  // padded: the device loop is unrolled by 4
    L.packed.assign((size_t)L.n_steps * T, -1);
    L.y.assign((size_t)L.n_steps * T, 0.0);
    L.run_base.assign(T, 0);
    L.run_start.assign(n_county + 1, 0);
    std::vector<int> run_county;
    for (int t = 0; t < T; ++t) {
        L.run_base[t] = (int32_t)run_county.size();
        int cur = -1;
        for (int j = 0; j < per; ++j) {
            const int pos = t * per + j;
            if (pos >= n_obs) break;
            const int o = order[pos];
            const int c = county[o];
            L.packed[(size_t)j * T + t] = (c << 1) | (floor[o] ? 1 : 0);
            L.y[(size_t)j * T + t] = y[o];
            if (c != cur) {
                run_county.push_back(c);
                cur = c;
            }
        }
    }
    L.R = (int)run_county.size();
    // runs are ordered by county because the observations are; county c owns
    // the contiguous run range [run_start[c], run_start[c+1])
    size_t r = 0;
    for (int c = 0; c <= n_county; ++c) {
        while (r < run_county.size() && run_county[r] < c) ++r;
        L.run_start[c] = (int32_t)r;
    }
    if (L.R == 0) L.R = 1;
    return L;
}

}  // namespace nb200
