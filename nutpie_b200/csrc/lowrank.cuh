// lowrank.cuh — the low-rank modified mass matrix (adaptation = "low_rank",
// PyNutsSettings::LowRank, src/wrapper.rs:307-346; python/nutpie/sample.py:921-933; the
// arithmetic is nuts-rs' mass_matrix/low_rank.rs, restated in oracle/lowrank.c), executed by the
// thread group that owns a chain, INSIDE the persistent sampler kernel.
//
//   M^-1 = S (I + V (L - I) V^T) S        S = diag(stds) [D], V [k][Dp] orthonormal rows,
//                                          L = diag(vals) [k]
//
// Per leapfrog the metric costs two contractions with V (k x D each way, V streamed from L2).
// The refresh — every mass_matrix_update_freq draws while the windows are open — estimates
//   Sigma = (X~X~^T + gamma I) # (G~G~^T + gamma I)^-1      (matrix geometric mean)
// from the chain's window of draws and gradients and keeps its eigenpairs beyond the cutoff.
// nuts-rs takes three symmetric eigen-decompositions and four matrix square roots through faer;
// here the geometric mean is taken through Cholesky factors, which needs only two matrices of
// scratch per chain and two one-sided (Hestenes) Jacobi runs whose inner loops are contiguous
// column sweeps — coalesced for a warp, and no eigenvector accumulation:
//   B = G~G~^T + gamma I = L L^T                      (Cholesky, in place)
//   M = L^T (X~X~^T + gamma I) L = C C^T              (two triangular products + Cholesky, in place)
//   C J1 = U Theta^1/2                                (one-sided Jacobi: M = U Theta U^T)
//   H = L^-T U Theta^1/4   =>   H H^T = L^-T M^1/2 L^-1 = Sigma      (back substitution)
//   H J2 = W Lambda^1/2                               (one-sided Jacobi: Sigma = W Lambda W^T)
// Written against the group policy G (tid / size / sync / reduce) like the rest of the core, so
// tests/emul runs the very same code on the CPU against the oracle's dense evaluation.
#pragma once
#include "group.cuh"
#include "portable.cuh"

namespace nb200 {

// per-chain low-rank state: pointers into the sampler's global buffers + the window counters
struct LrState {
    double* stds;   // [Dp]
    double* vals;   // [max_rank]
    double* vecs;   // [max_rank][Dp]
    double* coef;   // [max_rank] contraction coefficients (scratch)
    double* win;    // [cap][2][Dp] ring of (draw, gradient); logical entry j = slot (head + j) % cap
    double* matL;   // [D][Dp] column-major scratch (ld = Dp)
    double* matW;   // [D][Dp]
    double* cols;   // [6][Dp] scratch: mean x, mean g, scale x, scale g, key, lambda
    int k;          // eigenpairs in use
    int len, split, head, cap;  // window: [foreground-only part | background part] split at `split`
    int max_rank;
};

// coef_j = f_j * sum_i V_ji (sc_i x_i)   for all j < k   (sc = stds or nullptr for 1)
// mode 0: f_j = vals_j - 1 (velocity);  mode 1: f_j = 1 / sqrt(vals_j) - 1 (momentum)
template <class G>
NB_HD void lr_coefficients(const G& g, const LrState& L, int D, int Dp, const double* x,
                           const double* sc, int mode) {
    for (int j0 = 0; j0 < L.k; j0 += 8) {
        double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int i = g.tid; i < D; i += g.size()) {
            const double t = sc ? sc[i] * x[i] : x[i];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j0 + j < L.k) acc[j] += L.vecs[(size_t)(j0 + j) * Dp + i] * t;
        }
        g.reduce(acc);
        if (g.tid == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j0 + j < L.k) {
                    const double lam = L.vals[j0 + j];
                    L.coef[j0 + j] = (mode == 0 ? lam - 1.0 : 1.0 / sqrt(lam) - 1.0) * acc[j];
                }
        }
    }
    g.sync();
}

// v = M^-1 p   (p, v: [D]; may not alias)
template <class G>
NB_HD void lr_velocity(const G& g, const LrState& L, int D, int Dp, const double* p, double* v) {
    lr_coefficients(g, L, D, Dp, p, L.stds, 0);
    for (int i = g.tid; i < D; i += g.size()) {
        const double s = L.stds[i];
        double a = s * p[i];
        for (int j = 0; j < L.k; ++j) a += L.vecs[(size_t)j * Dp + i] * L.coef[j];
        v[i] = a * s;
    }
    g.sync();  // coef may be overwritten by the next contraction
}

// p <- M^1/2 z in place (z standard normal on entry), with the square root S^-1 (I + V (L^-1/2 - I) V^T)
template <class G>
NB_HD void lr_momentum(const G& g, const LrState& L, int D, int Dp, double* p) {
    lr_coefficients(g, L, D, Dp, p, (const double*)nullptr, 1);
    for (int i = g.tid; i < D; i += g.size()) {
        double a = p[i];
        for (int j = 0; j < L.k; ++j) a += L.vecs[(size_t)j * Dp + i] * L.coef[j];
        p[i] = a / L.stds[i];
    }
    g.sync();
}

// window deque -----------------------------------------------------------------------------
NB_HD double* lr_win_entry(const LrState& L, int Dp, int j, int which) {
    int slot = L.head + j;
    if (slot >= L.cap) slot -= L.cap;
    return L.win + ((size_t)slot * 2 + which) * Dp;
}
// append (q, grad) — grad_of(i) yields the gradient (stored or recomputed by the caller)
template <class G, class F>
NB_HD void lr_push(const G& g, LrState& L, int D, int Dp, const double* q, F&& grad_of) {
    if (L.len == L.cap) {  // cannot happen with cap = 3 * max switch_freq + 2; drop the oldest
        L.head = L.head + 1 == L.cap ? 0 : L.head + 1;
        L.len -= 1;
        if (L.split > 0) L.split -= 1;
    }
    double* wq = lr_win_entry(L, Dp, L.len, 0);
    double* wg = lr_win_entry(L, Dp, L.len, 1);
    for (int i = g.tid; i < D; i += g.size()) {
        wq[i] = q[i];
        wg[i] = grad_of(i);
    }
    L.len += 1;
    g.sync();
}
// the foreground-only part is dropped, what was the background becomes the foreground
NB_HD void lr_switch(LrState& L) {
    L.head = (L.head + L.split) % L.cap;
    L.len -= L.split;
    L.split = L.len;
}

// one-sided Jacobi on the nc columns (r rows each) of A (column-major, leading dimension ld): on
// return the columns are mutually orthogonal, A_out = A_in J with J orthogonal.  Round-robin (tournament) sweeps: a round
// pairs every column with exactly one other, so its pairs are independent — kJacobiPairs of them
// are rotated together (their column loads are in flight at the same time and their dot products
// share one reduction), which is what hides the L2 / HBM latency of a single warp walking a
// matrix that lives in global memory.  Sweeps until a sweep met no cosine above 1e-7 (the method
// converges quadratically: what that sweep left is below 1e-13, no verification sweep is run).
constexpr int kJacobiPairs = 4;
NB_HD double nb_rsqrt(double x) {
#ifdef __CUDA_ARCH__
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}
// Jacobi rotation that makes two columns with squared norms alpha, beta and inner product gam
// orthogonal: (c, s).  Returns false when they already are (|cos| <= 1e-13); `coarse` is set when
// the cosine was above 1e-7 (another sweep is needed after this one).  t = tan(theta) is the
// smaller root of t^2 + 2 zeta t - 1 = 0, zeta = (beta - alpha) / (2 gam), written on reciprocal
// square roots only: an IEEE division or square root is ~30 dependent instructions in fp64, and
// the four rotations of a step are most of its scalar work.
NB_HD bool lr_rotation(double alpha, double beta, double gam, double& c, double& s, bool& coarse) {
    const double ab = alpha * beta, g2 = gam * gam;
    if (!(g2 > 1e-26 * ab)) return false;
    coarse = coarse || g2 > 1e-14 * ab;
    const double a = beta - alpha, b = 2.0 * gam;
    const double d = a * a + b * b;
    const double h = d * nb_rsqrt(d);           // sqrt(a^2 + b^2) > 0
    const double rd = nb_rsqrt(fabs(a) + h);
    const double t = (a >= 0.0 ? b : -b) * (rd * rd);
    c = nb_rsqrt(1.0 + t * t);
    s = c * t;
    return true;
}
// pair k of round t in a tournament of m players (circle method; t < m - 1, k < m / 2)
NB_HD void lr_round_robin_pair(int t, int k, int m, int& a, int& b) {
    const int w = m - 1;  // t + k and t - k + w are below 2 w: one conditional subtraction each
    int x = t + k, y = t + w - k;
    x = x >= w ? x - w : x;
    y = y >= w ? y - w : y;
    a = k == 0 ? w : x;
    b = k == 0 ? t : y;
}
#ifdef __CUDACC__
// A warp per chain and r <= 32 NR: the columns of the kJacobiPairs pairs of a step live in
// REGISTERS (NR rows per lane) between the dot products and the rotation — one round trip to
// L2 / HBM per step with all its loads in flight together, instead of one per 32 rows and pass.
// Same arithmetic in the same order as the generic two-pass loop below.
template <int NR>
__device__ __noinline__ void lr_jacobi_warp(double* A, int r, int nc, int ld) {
    constexpr int NP = kJacobiPairs;
    const int lane = threadIdx.x & 31;
    const int m = (nc + 1) & ~1;
    for (int sweep = 0; sweep < 40; ++sweep) {
        int rotated = 0;
        bool coarse = false;  // a cosine above 1e-7 was rotated away in this sweep
        for (int t = 0; t < m - 1; ++t) {
            for (int k0 = 0; k0 < m / 2; k0 += NP) {
                double* ap[NP];
                double* aq[NP];
                bool on[NP];
                double x[NP][NR], y[NP][NR];
                double acc[3 * NP];
#pragma unroll
                for (int u = 0; u < NP; ++u) {
                    int a, b;
                    lr_round_robin_pair(t, k0 + u, m, a, b);
                    on[u] = k0 + u < m / 2 && a < nc && b < nc;
                    if (!on[u]) a = b = 0;
                    ap[u] = A + (size_t)a * ld;
                    aq[u] = A + (size_t)b * ld;
#pragma unroll
                    for (int it = 0; it < NR; ++it) {
                        const int i = lane + 32 * it;
                        x[u][it] = i < r ? ap[u][i] : 0.0;
                        y[u][it] = i < r ? aq[u][i] : 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < NP; ++u) {
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
                    for (int it = 0; it < NR; ++it) {
                        s0 += x[u][it] * x[u][it];
                        s1 += y[u][it] * y[u][it];
                        s2 += x[u][it] * y[u][it];
                    }
                    acc[3 * u] = s0;
                    acc[3 * u + 1] = s1;
                    acc[3 * u + 2] = s2;
                }
                GroupCuda<1>::warp_reduce(acc);
                bool any = false;
#pragma unroll
                for (int u = 0; u < NP; ++u) {
                    double c, sn;
                    if (on[u] && lr_rotation(acc[3 * u], acc[3 * u + 1], acc[3 * u + 2], c, sn, coarse)) {
#pragma unroll
                        for (int it = 0; it < NR; ++it) {
                            const int i = lane + 32 * it;
                            if (i < r) {
                                ap[u][i] = c * x[u][it] - sn * y[u][it];
                                aq[u][i] = sn * x[u][it] + c * y[u][it];
                            }
                        }
                        any = true;
                    }
                }
                if (any) {
                    rotated = 1;
                    __syncwarp();
                }
            }
        }
        // quadratic convergence: a sweep whose worst cosine was below 1e-7 leaves them below ~1e-13
#ifdef NB200_LR_PROFILE
        if (blockIdx.x == 0 && threadIdx.x == 0) printf("  jacobi sweep %d: coarse %d\n", sweep, (int)coarse);
#endif
        if (!rotated || !coarse) break;
    }
}
#endif
template <class G>
NB_HD void lr_jacobi_columns(const G& g, double* A, int r, int nc, int ld) {
    constexpr int NP = kJacobiPairs;
    if (nc < 2) return;
#ifdef __CUDA_ARCH__
    if constexpr (G::kThreads == 32) {
        switch ((r + 31) / 32) {
        case 1: lr_jacobi_warp<1>(A, r, nc, ld); return;
        case 2: lr_jacobi_warp<2>(A, r, nc, ld); return;
        case 3: lr_jacobi_warp<3>(A, r, nc, ld); return;
        case 4: lr_jacobi_warp<4>(A, r, nc, ld); return;
        case 5: case 6: lr_jacobi_warp<6>(A, r, nc, ld); return;
        case 7: case 8: lr_jacobi_warp<8>(A, r, nc, ld); return;
        default: break;  // wider matrices: the two-pass loop below
        }
    }
#endif
    const int m = (nc + 1) & ~1;  // players of the tournament (an odd count gets a bye: index nc)
    for (int sweep = 0; sweep < 40; ++sweep) {
        int rotated = 0;
        bool coarse = false;  // a cosine above 1e-7 was rotated away in this sweep
        for (int t = 0; t < m - 1; ++t) {
            for (int k0 = 0; k0 < m / 2; k0 += NP) {
                double* ap[NP];
                double* aq[NP];
                bool on[NP];
#pragma unroll
                for (int u = 0; u < NP; ++u) {
                    const int k = k0 + u;
                    int a, b;
                    lr_round_robin_pair(t, k, m, a, b);
                    on[u] = k < m / 2 && a < nc && b < nc;
                    if (!on[u]) a = b = 0;
                    ap[u] = A + (size_t)a * ld;
                    aq[u] = A + (size_t)b * ld;
                }
                double acc[3 * NP];
#pragma unroll
                for (int u = 0; u < 3 * NP; ++u) acc[u] = 0.0;
                for (int i = g.tid; i < r; i += g.size()) {
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        const double x = ap[u][i], y = aq[u][i];
                        acc[3 * u] += x * x;
                        acc[3 * u + 1] += y * y;
                        acc[3 * u + 2] += x * y;
                    }
                }
                g.reduce(acc);
                double c[NP], s[NP];
                bool any = false;
#pragma unroll
                for (int u = 0; u < NP; ++u) {
                    // (uniform over the group: the reduced sums are bit-identical on every thread)
                    if (on[u] && lr_rotation(acc[3 * u], acc[3 * u + 1], acc[3 * u + 2], c[u], s[u], coarse)) {
                        any = true;
                    } else {
                        on[u] = false;
                        c[u] = 1.0;
                        s[u] = 0.0;
                    }
                }
                if (!any) continue;
                rotated = 1;
                for (int i = g.tid; i < r; i += g.size()) {
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        if (on[u]) {
                            const double x = ap[u][i], y = aq[u][i];
                            ap[u][i] = c[u] * x - s[u] * y;
                            aq[u][i] = s[u] * x + c[u] * y;
                        }
                    }
                }
                g.sync();
            }
        }
        // quadratic convergence: a sweep whose worst cosine was below 1e-7 leaves them below ~1e-13
        if (!rotated || !coarse) break;
    }
}

// in-place Cholesky of the lower triangle of A [r][ld] (column-major): A = C C^T, the strict
// upper triangle is zeroed.  Returns false on a non-positive pivot.
template <class G>
NB_HD bool lr_cholesky(const G& g, double* A, int r, int ld) {
    for (int j = 0; j < r; ++j) {
        double* aj = A + (size_t)j * ld;
        for (int i = j + g.tid; i < r; i += g.size()) {
            double t = aj[i];
            for (int k = 0; k < j; ++k) t -= A[(size_t)k * ld + i] * A[(size_t)k * ld + j];
            aj[i] = t;
        }
        g.sync();
        const double piv = aj[j];
        g.sync();
        if (!(piv > 0.0) || !nb_isfinite(piv)) return false;
        const double d = sqrt(piv);
        for (int i = g.tid; i < r; i += g.size()) {
            if (i < j) aj[i] = 0.0;
            else if (i == j) aj[i] = d;
            else aj[i] = aj[i] / d;
        }
        g.sync();
    }
    return true;
}

// Refresh the metric from the window.  Returns false (metric unchanged) when the window is too
// short or a factorisation breaks down.
#if defined(NB200_LR_PROFILE) && defined(__CUDA_ARCH__)
#define NB_LR_MARK(name)                                                                       \
    do {                                                                                       \
        const long long now_ = clock64();                                                      \
        if (g.tid == 0 && blockIdx.x == 0 && threadIdx.x == 0)                                 \
            printf("lr_update n=%d r=%d %-12s %10.3f Mcycles\n", n, r, name, (now_ - mark_) * 1e-6); \
        mark_ = clock64();                                                                     \
    } while (0)
#define NB_LR_MARK_INIT() long long mark_ = clock64()
#else
#define NB_LR_MARK(name)
#define NB_LR_MARK_INIT()
#endif
// Sigma = A # B^-1 for SPD A (in W) and B (in Lm), both [r][ld] column-major, through Cholesky
// factors; on return the columns of W are the eigenvectors of Sigma scaled by the square roots of
// their eigenvalues (so lambda_j = |column j|^2).  Lm and W are destroyed.  False on breakdown.
template <class G>
NB_HD bool lr_geometric_mean_eig(const G& g, double* Lm, double* W, int r, int ld) {
    // Lm = chol(B)
    if (!lr_cholesky(g, Lm, r, ld)) return false;
    // W <- A Lm (ascending columns, in place), then W <- Lm^T W (lower triangle, in place)
    for (int c = 0; c < r; ++c) {
        for (int i = g.tid; i < r; i += g.size()) {
            double t = 0.0;
            for (int k = c; k < r; ++k) t += W[(size_t)k * ld + i] * Lm[(size_t)c * ld + k];
            W[(size_t)c * ld + i] = t;
        }
    }
    g.sync();
    for (int j = 0; j < r; ++j) {
        double* wj = W + (size_t)j * ld;
        for (int i0 = j; i0 < r; i0 += 8) {
            double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            for (int k = i0 + g.tid; k < r; k += g.size()) {
                const double t = wj[k];
#pragma unroll
                for (int m = 0; m < 8; ++m)
                    if (i0 + m < r && k >= i0 + m) acc[m] += Lm[(size_t)(i0 + m) * ld + k] * t;
            }
            g.reduce(acc);
            g.sync();  // every thread has read column j before its head is overwritten
            if (g.tid == 0) {
#pragma unroll
                for (int m = 0; m < 8; ++m)
                    if (i0 + m < r) wj[i0 + m] = acc[m];
            }
            g.sync();
        }
    }
    // W = chol(M), M = Lm^T A Lm
    if (!lr_cholesky(g, W, r, ld)) return false;
    // columns of W -> U Theta^1/2
    lr_jacobi_columns(g, W, r, r, ld);
    // W <- Lm^-T (W Theta^-1/4): every thread back-substitutes whole columns
    for (int j = g.tid; j < r; j += g.size()) {
        double* wj = W + (size_t)j * ld;
        double th = 0.0;
        for (int i = 0; i < r; ++i) th += wj[i] * wj[i];
        const double sc = th > 0.0 ? 1.0 / sqrt(sqrt(th)) : 0.0;
        for (int i = r - 1; i >= 0; --i) {
            double t = wj[i] * sc;
            const double* li = Lm + (size_t)i * ld;
            for (int k = i + 1; k < r; ++k) t -= li[k] * wj[k];
            wj[i] = t / li[i];
        }
    }
    g.sync();
    // columns of W -> W_sigma Lambda^1/2
    lr_jacobi_columns(g, W, r, r, ld);
    return true;
}

// Refresh the metric from the window.  Returns false (metric unchanged) when the window is too
// short or a factorisation breaks down.
//   * D <= 10 n / 3: the matrices are formed in the full space (r = D).
//   * fewer draws: both covariances are gamma I outside the span of the 2 n window vectors and
//     Sigma is the identity there (eigenvalue 1, never kept), so the estimate is taken in that
//     span — an orthonormal basis Q [D x r], r <= 2 n, from a one-sided Jacobi run on the window
//     vectors themselves, the r x r problem in Q's coordinates, eigenvectors mapped back through
//     Q (what nuts-rs does with thin SVDs and a pivoted QR).  The early windows hold 10-20 draws:
//     the refresh is then a 40 x 40 problem instead of a D x D one.
template <class G>
NB_HD bool lr_update(const G& g, LrState& L, int D, int Dp, double gamma, double cutoff) {
    const int n = L.len;
    if (n < 3) return false;
    int r = D, ld = Dp;          // size / leading dimension of the eigenproblem
    NB_LR_MARK_INIT();
    double* mx = L.cols;
    double* mg = L.cols + Dp;
    double* xs = L.cols + 2 * (size_t)Dp;
    double* gs = L.cols + 3 * (size_t)Dp;
    double* key = L.cols + 4 * (size_t)Dp;
    double* lam = L.cols + 5 * (size_t)Dp;
    const double dn = (double)n;
    // ---- per-dimension means, scales: stds_i = sqrt(sd(x_i) / sd(g_i))
    for (int i = g.tid; i < D; i += g.size()) {
        double sx = 0.0, sg = 0.0;
        for (int j = 0; j < n; ++j) {
            sx += lr_win_entry(L, Dp, j, 0)[i];
            sg += lr_win_entry(L, Dp, j, 1)[i];
        }
        sx /= dn;
        sg /= dn;
        double vx = 0.0, vg = 0.0;
        for (int j = 0; j < n; ++j) {
            const double a = lr_win_entry(L, Dp, j, 0)[i] - sx, b = lr_win_entry(L, Dp, j, 1)[i] - sg;
            vx += a * a;
            vg += b * b;
        }
        double s = sqrt(sqrt(vx / dn) / sqrt(vg / dn));
        if (!nb_isfinite(s) || s <= 0.0) s = L.stds[i];
        if (s < 1e-10) s = 1e-10;
        if (s > 1e10) s = 1e10;
        L.stds[i] = s;
        mx[i] = sx;
        mg[i] = sg;
        xs[i] = 1.0 / (s * sqrt(dn));
        gs[i] = s / sqrt(dn);
    }
    g.sync();
    NB_LR_MARK("stats");
    // (the subspace form must fit the two D x Dp scratch matrices: Q and the projections in one,
    // the two small matrices in the other)
    const bool subspace = 10 * n <= 3 * D &&
                          (size_t)2 * n * Dp + (size_t)2 * (2 * n + 3) * n <= (size_t)D * Dp &&
                          (size_t)2 * (2 * n + 3) * 2 * n <= (size_t)D * Dp;
    double* Lm = L.matL;         // full space: B, then its Cholesky factor
    double* W = L.matW;          // full space: A, ..., eigenvectors of Sigma
    const double* Q = nullptr;   // subspace: orthonormal basis [D][Dp] (first r columns of matL)
    if (!subspace) {
        // ---- W = X~X~^T + gamma I,  Lm = G~G~^T + gamma I   (full symmetric storage)
        for (int a0 = 0; a0 < r; a0 += g.size()) {
            const int a = a0 + g.tid;
            const bool live = a < r;
            const double mxa = live ? mx[a] : 0.0, mga = live ? mg[a] : 0.0;
            const double xsa = live ? xs[a] : 0.0, gsa = live ? gs[a] : 0.0;
            const int bmax = a0 + g.size() < r ? a0 + g.size() : r;  // columns b <= the chunk's last row
            for (int b = 0; b < bmax; ++b) {
                const double mxb = mx[b], mgb = mg[b], xsb = xs[b], gsb = gs[b];
                double ax = 0.0, ag = 0.0;
                if (live && b <= a) {
                    for (int j = 0; j < n; ++j) {
                        const double* wq = lr_win_entry(L, Dp, j, 0);
                        const double* wg = lr_win_entry(L, Dp, j, 1);
                        ax += ((wq[a] - mxa) * xsa) * ((wq[b] - mxb) * xsb);
                        ag += ((wg[a] - mga) * gsa) * ((wg[b] - mgb) * gsb);
                    }
                    if (a == b) {
                        ax += gamma;
                        ag += gamma;
                    }
                    W[(size_t)b * ld + a] = ax;
                    W[(size_t)a * ld + b] = ax;
                    Lm[(size_t)b * ld + a] = ag;
                    Lm[(size_t)a * ld + b] = ag;
                }
            }
        }
        g.sync();
        NB_LR_MARK("gram");
    } else {
        // ---- Z = [X~ G~] in matL (2 n columns of D rows); its columns orthogonalised in place
        double* Z = L.matL;
        const int nz = 2 * n;
        for (int c = 0; c < nz; ++c) {
            const double* w = lr_win_entry(L, Dp, c < n ? c : c - n, c < n ? 0 : 1);
            const double* mean = c < n ? mx : mg;
            const double* scale = c < n ? xs : gs;
            double* zc = Z + (size_t)c * Dp;
            for (int i = g.tid; i < D; i += g.size()) zc[i] = (w[i] - mean[i]) * scale[i];
        }
        g.sync();
        lr_jacobi_columns(g, Z, D, nz, Dp);
        // squared norms of the orthogonal columns; the ones above the noise floor are kept,
        // normalised and packed to the front: Q
        double smax = 0.0;
        for (int c0 = 0; c0 < nz; c0 += 8) {
            double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            for (int i = g.tid; i < D; i += g.size()) {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (c0 + u < nz) {
                        const double z = Z[(size_t)(c0 + u) * Dp + i];
                        acc[u] += z * z;
                    }
            }
            g.reduce(acc);
            g.sync();
            if (g.tid == 0) {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (c0 + u < nz) key[c0 + u] = acc[u];  // (key / lam are free until the selection)
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (c0 + u < nz && acc[u] > smax) smax = acc[u];
            g.sync();
        }
        if (!(smax > 0.0) || !nb_isfinite(smax)) return false;
        r = 0;
        for (int c = 0; c < nz; ++c) {
            const double s2 = key[c];
            if (!(s2 > 1e-12 * smax)) continue;  // uniform: every thread reads the same value
            const double inv = 1.0 / sqrt(s2);
            const double* src = Z + (size_t)c * Dp;
            double* dst = Z + (size_t)r * Dp;  // r <= c: an earlier or the same column
            for (int i = g.tid; i < D; i += g.size()) dst[i] = src[i] * inv;
            ++r;
        }
        g.sync();
        if (r < 1) return false;
        Q = Z;
        ld = (r + 3) & ~3;
        // small matrices in matW: B_p, A_p [r][ld]; projections X_p, G_p [n][ld] behind Q in matL
        Lm = L.matW;
        W = L.matW + (size_t)ld * r;
        double* Xp = L.matL + (size_t)r * Dp;
        double* Gp = Xp + (size_t)ld * n;
        for (int j = 0; j < n; ++j) {
            const double* wq = lr_win_entry(L, Dp, j, 0);
            const double* wg = lr_win_entry(L, Dp, j, 1);
            for (int a0 = 0; a0 < r; a0 += 4) {
                double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                for (int i = g.tid; i < D; i += g.size()) {
                    const double x = (wq[i] - mx[i]) * xs[i], y = (wg[i] - mg[i]) * gs[i];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (a0 + u < r) {
                            const double q = Q[(size_t)(a0 + u) * Dp + i];
                            acc[u] += q * x;
                            acc[4 + u] += q * y;
                        }
                }
                g.reduce(acc);
                if (g.tid == 0) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (a0 + u < r) {
                            Xp[(size_t)j * ld + a0 + u] = acc[u];
                            Gp[(size_t)j * ld + a0 + u] = acc[4 + u];
                        }
                }
            }
        }
        g.sync();
        // A_p = X_p X_p^T + gamma I -> W,  B_p = G_p G_p^T + gamma I -> Lm
        for (int a = g.tid; a < r; a += g.size()) {
            for (int b = 0; b < r; ++b) {
                double ax = 0.0, ag = 0.0;
                for (int j = 0; j < n; ++j) {
                    ax += Xp[(size_t)j * ld + a] * Xp[(size_t)j * ld + b];
                    ag += Gp[(size_t)j * ld + a] * Gp[(size_t)j * ld + b];
                }
                if (a == b) {
                    ax += gamma;
                    ag += gamma;
                }
                W[(size_t)b * ld + a] = ax;
                Lm[(size_t)b * ld + a] = ag;
            }
        }
        g.sync();
        NB_LR_MARK("subspace");
    }
    if (!lr_geometric_mean_eig(g, Lm, W, r, ld)) return false;
    NB_LR_MARK("eig");
    // ---- keep the eigenpairs beyond the cutoff, largest |log lambda| first
    for (int j = g.tid; j < r; j += g.size()) {
        const double* wj = W + (size_t)j * ld;
        double l2 = 0.0;
        for (int i = 0; i < r; ++i) l2 += wj[i] * wj[i];
        lam[j] = l2;
        const bool keep = nb_isfinite(l2) && l2 > 0.0 && (l2 > cutoff || l2 < 1.0 / cutoff);
        key[j] = keep ? fabs(log(l2)) : -1.0;
    }
    g.sync();
    int k = 0;
    const int kmax = L.max_rank < r ? L.max_rank : r;
    while (k < kmax) {
        int best = -1;
        double bk = -1.0;  // keepers have key > 0, everything else -1
        for (int j = 0; j < r; ++j) {  // every thread scans the same keys: uniform result
            const double kj = key[j];
            if (kj > bk) {
                best = j;
                bk = kj;
            }
        }
        if (best < 0) break;
        const double l2 = lam[best];
        const double inv = 1.0 / sqrt(l2);
        const double* wb = W + (size_t)best * ld;
        double* vk = L.vecs + (size_t)k * Dp;
        if (Q) {  // back to the full space: v = Q w
            for (int i = g.tid; i < D; i += g.size()) {
                double a = 0.0;
                for (int c = 0; c < r; ++c) a += Q[(size_t)c * Dp + i] * wb[c];
                vk[i] = a * inv;
            }
        } else {
            for (int i = g.tid; i < D; i += g.size()) vk[i] = wb[i] * inv;
        }
        g.sync();  // everyone has read key[] before it changes
        if (g.tid == 0) {
            L.vals[k] = l2;
            key[best] = -1.0;
        }
        g.sync();
        ++k;
    }
    L.k = k;
    NB_LR_MARK("select");
    return true;
}

}  // namespace nb200
