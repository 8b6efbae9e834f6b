// philox.cuh — counter-based random stream of the B200 NUTS engine.
//
// Philox4x32-10, key = 64-bit seed, counter = (index, purpose, draw, chain).
// Replaces the per-chain ChaCha8 stream nuts-rs derives with
// `rng.set_stream(chain + 1)` (nutpie consumes it at
// src/stan.rs:788,803-806, src/pymc.rs:510, src/pyfunc.rs:541-542).  Keying by
// the GLOBAL chain id makes a run independent of how chains are sharded over
// GPUs.
#pragma once
#include <stdint.h>
#include "portable.cuh"

namespace nb200 {

enum : uint32_t {
    RNG_MOMENTUM = 0,   // index = element pair j -> z[2j], z[2j+1]
    RNG_DIRECTION = 1,  // index = tree depth at the doubling
    RNG_MERGE = 2,      // index = merge sequence number inside the draw
    RNG_INIT_POS = 3,   // draw = attempt number, index = element pair
    RNG_STEP_INIT = 4,  // draw = draw index (0xFFFFFFFF before the first draw)
    RNG_JITTER = 5      // index 0: step-size jitter factor of the draw
};

NB_HD void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                                uint32_t c3, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

NB_HD void rng_u64x2(uint64_t seed, uint32_t chain, uint32_t draw, uint32_t purpose,
                            uint32_t index, uint64_t& a, uint64_t& b) {
    uint32_t o[4];
    philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), index, purpose, draw, chain, o);
    a = (uint64_t)o[0] | ((uint64_t)o[1] << 32);
    b = (uint64_t)o[2] | ((uint64_t)o[3] << 32);
}

NB_HD double rng_u01(uint64_t x) { return (double)(x >> 11) * 0x1.0p-53; }
NB_HD double rng_u01_open0(uint64_t x) { return ((double)(x >> 11) + 1.0) * 0x1.0p-53; }

// Box–Muller pair
NB_HD void rng_normal_pair(uint64_t a, uint64_t b, double& z0, double& z1) {
    double u1 = rng_u01_open0(a), u2 = rng_u01(b);
    double r = sqrt(-2.0 * log(u1));
    double s, c;
    nb_sincos_2pi(u2, s, c);
    z0 = r * c;
    z1 = r * s;
}

}  // namespace nb200
