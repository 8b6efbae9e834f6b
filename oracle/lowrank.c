/*
 * oracle/lowrank.c — TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement of nuts-rs' low-rank modified mass matrix (adaptation="low_rank":
 * src/wrapper.rs:307-346, python/nutpie/sample.py:921-933, docs/sampling-options.qmd:124-144;
 * the arithmetic is in the un-vendored crate nuts-rs 0.18.3, mass_matrix/low_rank.rs,
 * restated here from its published description — PARITY UNPINNED like the rest of oracle/).
 *
 * The metric:   M^-1 = S (I + V (L - I) V^T) S     S = diag(stds), V [dim x k] orthonormal,
 *                                                   L = diag(vals)
 * The estimate from a window of n draws x_j and gradients g_j:
 *   stds_i  = sqrt( sd(x_i) / sd(g_i) )                       (population sd over the window)
 *   X~      = (x - mean x) / (stds sqrt n),   G~ = (g - mean g) stds / sqrt n
 *   Cx      = X~ X~^T + gamma I,   Cg = G~ G~^T + gamma I      (regularised covariances)
 *   Sigma   = Cx # Cg^-1 = Cg^-1/2 (Cg^1/2 Cx Cg^1/2)^1/2 Cg^-1/2   (geometric mean: the
 *             metric that turns the gradient covariance into the inverse draw covariance)
 *   (vals, V) = eigenpairs of Sigma with vals > cutoff or vals < 1 / cutoff
 * nuts-rs evaluates this in the span of [X~ G~] (thin SVDs + a pivoted QR); outside that span
 * both covariances are gamma I and Sigma is the identity (eigenvalue 1, never kept), so the
 * full-space evaluation below gives the same eigenpairs.  Every matrix function goes through
 * a cyclic two-sided Jacobi eigensolver — slow, short and hard to get wrong; the CUDA engine
 * uses a different route (Cholesky factors + one-sided Jacobi) and is compared with this one
 * as an OPERATOR (tests/test_lowrank.py).
 * Ours, not nuts-rs': at most max_rank eigenpairs are kept (largest |log val| first).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

/* A [n][n] symmetric (destroyed), V [n][n]: column j = eigenvector j, w [n] eigenvalues */
static void jacobi_eigh(int n, double *A, double *V, double *w) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V[i * n + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; ++i) {
            diag += A[i * n + i] * A[i * n + i];
            for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j];
        }
        if (off <= 1e-30 * diag || off == 0.0) break;
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                double apq = A[p * n + q];
                if (apq == 0.0) continue;
                double app = A[p * n + p], aqq = A[q * n + q];
                double theta = (aqq - app) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) { /* columns p, q */
                    double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) { /* rows p, q */
                    double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
        }
    }
    for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
}

/* out = V diag(f) V^T */
static void recompose(int n, const double *V, const double *f, double *out) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) s += V[i * n + k] * f[k] * V[j * n + k];
            out[i * n + j] = out[j * n + i] = s;
        }
}
/* out = A B (all [n][n]) */
static void matmul(int n, const double *A, const double *B, double *out) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) s += A[i * n + k] * B[k * n + j];
            out[i * n + j] = s;
        }
}
static void symmetrise(int n, double *A) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) {
            double m = 0.5 * (A[i * n + j] + A[j * n + i]);
            A[i * n + j] = A[j * n + i] = m;
        }
}

/* geometric mean  A # B^-1  of two SPD matrices (nuts-rs spd_mean); A, B are destroyed */
static void spd_mean(int n, double *A, double *B, double *out) {
    size_t nn = (size_t)n * n;
    double *V = malloc(nn * 8), *w = malloc((size_t)n * 8), *f = malloc((size_t)n * 8);
    double *Bh = malloc(nn * 8), *Bmh = malloc(nn * 8), *T = malloc(nn * 8), *Mm = malloc(nn * 8);
    jacobi_eigh(n, B, V, w);
    for (int i = 0; i < n; ++i) f[i] = sqrt(w[i]);
    recompose(n, V, f, Bh);
    for (int i = 0; i < n; ++i) f[i] = 1.0 / sqrt(w[i]);
    recompose(n, V, f, Bmh);
    matmul(n, Bh, A, T);
    matmul(n, T, Bh, Mm);
    symmetrise(n, Mm);
    jacobi_eigh(n, Mm, V, w);
    for (int i = 0; i < n; ++i) f[i] = sqrt(w[i] > 0 ? w[i] : 0.0);
    recompose(n, V, f, Mm);
    matmul(n, Bmh, Mm, T);
    matmul(n, T, Bmh, out);
    symmetrise(n, out);
    free(V); free(w); free(f); free(Bh); free(Bmh); free(T); free(Mm);
}

typedef struct { double key; int idx; } Ranked;
static int ranked_cmp(const void *a, const void *b) {
    double ka = ((const Ranked *)a)->key, kb = ((const Ranked *)b)->key;
    if (ka != kb) return ka > kb ? -1 : 1;
    return ((const Ranked *)a)->idx - ((const Ranked *)b)->idx;
}

/* stds [dim] holds the previous scales on entry (kept where a window has no spread) */
int oracle_lowrank_update(size_t dim, size_t n, const double *draws, const double *grads,
                          double gamma, double cutoff, size_t max_rank, double *stds,
                          double *vals, double *vecs, size_t *rank_out) {
    *rank_out = 0;
    if (n < 2) return 1;
    int d = (int)dim;
    size_t dd = dim * dim;
    double *X = malloc(dim * n * 8), *G = malloc(dim * n * 8); /* [dim][n] scaled + centred */
    double *Cx = malloc(dd * 8), *Cg = malloc(dd * 8), *Sg = malloc(dd * 8);
    double *W = malloc(dd * 8), *lam = malloc(dim * 8);
    for (size_t i = 0; i < dim; ++i) {
        double mx = 0.0, mg = 0.0;
        for (size_t j = 0; j < n; ++j) { mx += draws[j * dim + i]; mg += grads[j * dim + i]; }
        mx /= (double)n; mg /= (double)n;
        double vx = 0.0, vg = 0.0;
        for (size_t j = 0; j < n; ++j) {
            double a = draws[j * dim + i] - mx, b = grads[j * dim + i] - mg;
            vx += a * a; vg += b * b;
        }
        double s = sqrt(sqrt(vx / (double)n) / sqrt(vg / (double)n));
        if (!isfinite(s) || s <= 0.0) s = stds[i];
        if (s < 1e-10) s = 1e-10;
        if (s > 1e10) s = 1e10;
        stds[i] = s;
        double xs = 1.0 / (s * sqrt((double)n)), gs = s / sqrt((double)n);
        for (size_t j = 0; j < n; ++j) {
            X[i * n + j] = (draws[j * dim + i] - mx) * xs;
            G[i * n + j] = (grads[j * dim + i] - mg) * gs;
        }
    }
    for (size_t a = 0; a < dim; ++a)
        for (size_t b = 0; b <= a; ++b) {
            double sx = 0.0, sg = 0.0;
            for (size_t j = 0; j < n; ++j) {
                sx += X[a * n + j] * X[b * n + j];
                sg += G[a * n + j] * G[b * n + j];
            }
            if (a == b) { sx += gamma; sg += gamma; }
            Cx[a * dim + b] = Cx[b * dim + a] = sx;
            Cg[a * dim + b] = Cg[b * dim + a] = sg;
        }
    spd_mean(d, Cx, Cg, Sg);
    jacobi_eigh(d, Sg, W, lam);
    Ranked *r = malloc(dim * sizeof(Ranked));
    size_t m = 0;
    for (size_t i = 0; i < dim; ++i)
        if (isfinite(lam[i]) && lam[i] > 0.0 && (lam[i] > cutoff || lam[i] < 1.0 / cutoff)) {
            r[m].key = fabs(log(lam[i]));
            r[m].idx = (int)i;
            ++m;
        }
    qsort(r, m, sizeof(Ranked), ranked_cmp);
    if (m > max_rank) m = max_rank;
    for (size_t k = 0; k < m; ++k) {
        vals[k] = lam[r[k].idx];
        for (size_t i = 0; i < dim; ++i) vecs[k * dim + i] = W[i * dim + (size_t)r[k].idx];
    }
    *rank_out = m;
    free(r); free(X); free(G); free(Cx); free(Cg); free(Sg); free(W); free(lam);
    return 0;
}

/* v = M^-1 p */
void oracle_lowrank_velocity(size_t dim, const double *stds, size_t k, const double *vals,
                             const double *vecs, const double *p, double *v) {
    for (size_t i = 0; i < dim; ++i) v[i] = stds[i] * p[i];
    if (k > 0) {
        double *c = malloc(k * 8);
        for (size_t j = 0; j < k; ++j) {
            double s = 0.0;
            for (size_t i = 0; i < dim; ++i) s += vecs[j * dim + i] * v[i];
            c[j] = (vals[j] - 1.0) * s;
        }
        for (size_t i = 0; i < dim; ++i) {
            double a = v[i];
            for (size_t j = 0; j < k; ++j) a += vecs[j * dim + i] * c[j];
            v[i] = a;
        }
        free(c);
    }
    for (size_t i = 0; i < dim; ++i) v[i] *= stds[i];
}

/* p = M^1/2 z with the square root  S^-1 (I + V (L^-1/2 - I) V^T) */
void oracle_lowrank_momentum(size_t dim, const double *stds, size_t k, const double *vals,
                             const double *vecs, const double *z, double *p) {
    for (size_t i = 0; i < dim; ++i) p[i] = z[i];
    if (k > 0) {
        double *c = malloc(k * 8);
        for (size_t j = 0; j < k; ++j) {
            double s = 0.0;
            for (size_t i = 0; i < dim; ++i) s += vecs[j * dim + i] * z[i];
            c[j] = (1.0 / sqrt(vals[j]) - 1.0) * s;
        }
        for (size_t i = 0; i < dim; ++i) {
            double a = z[i];
            for (size_t j = 0; j < k; ++j) a += vecs[j * dim + i] * c[j];
            p[i] = a;
        }
        free(c);
    }
    for (size_t i = 0; i < dim; ++i) p[i] /= stds[i];
}
