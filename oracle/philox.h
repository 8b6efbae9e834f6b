/*
 * oracle/philox.h — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Counter-based random numbers shared (by specification, not by source) with
 * the CUDA engine: Philox4x32-10 (Salmon et al., SC'11), key = 64-bit seed,
 * counter = (index, purpose, draw, chain).
 *
 * The reference draws from rand 0.10 ChaCha8 with stream = chain+1
 * (the call sites are src/stan.rs:788,803-806,
 * src/pymc.rs:510, src/pyfunc.rs:541-542).  Reproducing that bit stream is
 * out of reach without the rand/rand_distr sources, so parity with the
 * reference is statistical; parity between this oracle and the CUDA engine
 * is draw-for-draw because both use the stream defined here.
 */
#ifndef ORACLE_PHILOX_H
#define ORACLE_PHILOX_H
#include <math.h>
#include <stdint.h>

enum {
    RNG_MOMENTUM = 0,   /* index = element pair j -> z[2j], z[2j+1]            */
    RNG_DIRECTION = 1,  /* index = tree depth at the doubling                   */
    RNG_MERGE = 2,      /* index = merge sequence number inside the draw        */
    RNG_INIT_POS = 3,   /* draw = attempt number, index = element pair          */
    RNG_STEP_INIT = 4,  /* momentum of the step-size search; draw = draw index, */
                        /* 0xFFFFFFFF for the search before the first draw      */
    RNG_JITTER = 5      /* index 0: step-size jitter factor of the draw         */
};

static inline void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                 uint32_t c2, uint32_t c3, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n1 = lo1;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        uint32_t n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* two 64-bit words for (seed; index, purpose, draw, chain) */
static inline void rng_u64x2(uint64_t seed, uint32_t chain, uint32_t draw, uint32_t purpose,
                             uint32_t index, uint64_t *a, uint64_t *b) {
    uint32_t o[4];
    philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), index, purpose, draw, chain, o);
    *a = (uint64_t)o[0] | ((uint64_t)o[1] << 32);
    *b = (uint64_t)o[2] | ((uint64_t)o[3] << 32);
}

/* [0,1) with 53 bits */
static inline double rng_u01(uint64_t x) { return (double)(x >> 11) * 0x1.0p-53; }
/* (0,1] with 53 bits */
static inline double rng_u01_open0(uint64_t x) { return ((double)(x >> 11) + 1.0) * 0x1.0p-53; }

/* Box–Muller pair from two words */
static inline void rng_normal_pair(uint64_t a, uint64_t b, double *z0, double *z1) {
    double u1 = rng_u01_open0(a), u2 = rng_u01(b);
    double r = sqrt(-2.0 * log(u1));
    double t = 6.283185307179586476925286766559 * u2;
    *z0 = r * cos(t);
    *z1 = r * sin(t);
}
#endif
