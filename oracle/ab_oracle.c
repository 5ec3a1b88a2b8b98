/* ab_oracle.c — CPU restatement (plain C) of the reference's AlmostBandedMatrix QR + solve.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * What it restates (reference = /root/reference, SemiseparableMatrices.jl v0.4.1):
 *   ab_getindex        src/AlmostBandedMatrix.jl:84-90      A[k,j] = bands[k,j] if j <= k+u else (U*V)[k,j]
 *   ab_widen           src/AlmostBandedMatrix.jl:30-38      copy to bandwidths (l, l+u); new bands filled from U*V
 *   ab_qr_inplace      src/AlmostBandedMatrix.jl:135-172    blocked Householder sweep with tracked rank-r fill
 *   ab_lmul_qt         src/AlmostBandedMatrix.jl:181,189    b <- Q'b, Q = QRPackedQ(bandpart(factors), tau)
 *   ab_upper_ldiv      src/AlmostBandedMatrix.jl:215-240    blocked back-substitution with running r-vector
 *   ab_lower_ldiv      src/AlmostBandedMatrix.jl:258-270    lower-triangular solves use the band part only
 *   ab_solve           src/AlmostBandedMatrix.jl:125-133,187-192   A \ b = qr(A) then ldiv!
 *
 * Third-party arithmetic that is NOT vendored in /root/reference and is restated from its
 * published algorithm (dependency versions from Project.toml compat: BandedMatrices 1.7,
 * MatrixFactorizations 3, Julia >= 1.10 LinearAlgebra):
 *   reflector_         LinearAlgebra.reflector!      nu = copysign(||x||, x1); x1 <- -nu; x[2:] /= (x1+nu); tau = (x1+nu)/nu
 *   reflector_apply_   LinearAlgebra.reflectorApply! vAj = tau*(A[1,j] + x[2:]'A[2:,j]); A[1,j] -= vAj; A[2:,j] -= x[2:]*vAj
 *   banded_qr_block_   BandedMatrices._banded_qr!    column-by-column reflectors of length l+1 inside a banded block
 *   banded_lmul_qt_    BandedMatrices banded QRPackedQ' lmul!   vBj = B[k,j] + sum v_i B[i,j]; vBj *= tau; ...
 *
 * Parity pinning: the reference's own tests pin this path only through dense LinearAlgebra
 * (test/test_almostbanded.jl:145,153,163,176,184) and the README known-answer vector
 * (README.md:50,73-94).  tests/test_oracle_ab.py checks this file against LAPACK dgeqrf /
 * numpy.linalg.solve on re-instantiations of those tests, and against the README vector.
 * Julia is not available in this image, so the reference itself cannot be executed here.
 *
 * All matrices use the reference's memory layout.  Index arguments of the static helpers
 * are 1-based so the code reads side by side with the Julia source.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/ssm_rng.h"

typedef struct {
    int n, l, u, r;
    double *bands; /* (l+u+1) x n column-major band storage, A[k,j] at row u+1+k-j (1-based) of column j */
    double *U;     /* n x r column-major */
    double *V;     /* r x n column-major */
} ab_t;

static inline double *bptr(const ab_t *A, int k, int j) {
    return &A->bands[(size_t)(A->u + k - j) + (size_t)(A->l + A->u + 1) * (size_t)(j - 1)];
}
static inline int inband(const ab_t *A, int k, int j) { return (j - k <= A->u) && (k - j <= A->l); }
static inline double band_get(const ab_t *A, int k, int j) { return inband(A, k, j) ? *bptr(A, k, j) : 0.0; }
static inline double Uget(const ab_t *A, int k, int q) { return A->U[(size_t)(k - 1) + (size_t)A->n * (size_t)(q - 1)]; }
static inline double *Uptr(const ab_t *A, int k, int q) { return &A->U[(size_t)(k - 1) + (size_t)A->n * (size_t)(q - 1)]; }
static inline double Vget(const ab_t *A, int q, int j) { return A->V[(size_t)(q - 1) + (size_t)A->r * (size_t)(j - 1)]; }
/* LowRankMatrix = ApplyMatrix(*, U, V) (src/SemiseparableMatrices.jl:20-21): entry = sum_q U[k,q] V[q,j] */
static inline double fill_get(const ab_t *A, int k, int j) {
    double s = 0.0;
    for (int q = 1; q <= A->r; ++q) s += Uget(A, k, q) * Vget(A, q, j);
    return s;
}
/* src/AlmostBandedMatrix.jl:84-90 */
static inline double ab_getindex(const ab_t *A, int k, int j) {
    if (j > k + A->u) return fill_get(A, k, j);
    return band_get(A, k, j);
}

/* LinearAlgebra.generic_norm2 (dependency; small vectors never reach BLAS nrm2) */
static double norm2_(const double *x, int len) {
    double maxabs = 0.0;
    for (int i = 0; i < len; ++i) { double a = fabs(x[i]); if (a > maxabs || a != a) maxabs = a; }
    if (maxabs == 0.0 || isinf(maxabs) || maxabs != maxabs) return maxabs;
    double mm = maxabs * maxabs;
    if (isfinite((double)len * mm) && mm != 0.0) {
        double s = 0.0;
        for (int i = 0; i < len; ++i) s += x[i] * x[i];
        return sqrt(s);
    }
    double s = 0.0;
    for (int i = 0; i < len; ++i) { double t = x[i] / maxabs; s += t * t; }
    return maxabs * sqrt(s);
}

/* LinearAlgebra.reflector! on a contiguous vector */
static double reflector_(double *x, int len) {
    if (len == 0) return 0.0;
    double xi1 = x[0];
    double normu = norm2_(x, len);
    if (normu == 0.0) return 0.0;
    double nu = copysign(normu, xi1);
    xi1 += nu;
    x[0] = -nu;
    for (int i = 1; i < len; ++i) x[i] /= xi1;
    return xi1 / nu;
}

/* src/AlmostBandedMatrix.jl:30-38 — widened copy: bandwidths (l, l+u); bands u+1..l+u take the fill values. */
static void ab_widen(const ab_t *A, ab_t *W /* W->bands preallocated (2l+u+1) x n, W->U preallocated copy target */) {
    int n = A->n, l = A->l, u = A->u;
    W->n = n; W->l = l; W->u = l + u; W->r = A->r;
    memset(W->bands, 0, sizeof(double) * (size_t)(2 * l + u + 1) * (size_t)n);
    for (int j = 1; j <= n; ++j)                                   /* B = BandedMatrix(B_in, (l, l+u))  :33 */
        for (int k = (j - u > 1 ? j - u : 1); k <= (j + l < n ? j + l : n); ++k) *bptr(W, k, j) = *bptr(A, k, j);
    memcpy(W->U, A->U, sizeof(double) * (size_t)n * (size_t)A->r);  /* deepcopy(L)                       :37 */
    W->V = A->V;                                                    /* V is never written by the path       */
    for (int b = u + 1; b <= l + u; ++b)                            /* B[band(b)] .= L[band(b)]       :34-36 */
        for (int k = 1; k + b <= n; ++k) *bptr(W, k, k + b) = fill_get(A, k, k + b);
}

/* src/AlmostBandedMatrix.jl:135-172.  A holds the widened bands (l, u=ubar); tau has length n (zeros on entry). */
static void ab_qr_inplace(ab_t *A, double *tau, int ncols) {
    int n = A->n, m = A->n, l = A->l, u = A->u, r = A->r;
    int k = 1;
    while (k <= ncols) {
        int kr_end = (k + l + u < m) ? k + l + u : m;              /* kr  = k:min(k+l+u,m)        :145 */
        int jr1_end = (k + u < n) ? k + u : n;                     /* jr1 = k:min(k+u,n)          :146 */
        int jr2_beg = k + u + 1;                                   /* jr2 = k+u+1:min(kr[end]+u,n) :147 */
        int jr2_end = (kr_end + u < n) ? kr_end + u : n;
        int jr3_end = jr1_end < ncols ? jr1_end : ncols;           /* jr3 = k:min(k+u,n,ncols)    :148 */

        /* R,_ = _banded_qr!(S, tauv, length(jr3))  :152   (BandedMatrices, restated) */
        for (int kk = k; kk <= jr3_end; ++kk) {
            int re = (kk + l < kr_end) ? kk + l : kr_end;          /* rows kk..re of column kk          */
            double *x = bptr(A, kk, kk);                           /* contiguous in band storage        */
            int len = re - kk + 1;
            double t = reflector_(x, len);
            tau[kk - 1] = t;
            int jmax = (u < jr1_end - kk) ? u : jr1_end - kk;
            for (int j = 1; j <= jmax; ++j) {                      /* reflectorApply!(x, tau, column kk+j) */
                double *a = bptr(A, kk, kk + j);
                double dot = 0.0;
                for (int i = 1; i < len; ++i) dot += x[i] * a[i];
                double vAj = t * (a[0] + dot);
                a[0] -= vAj;
                for (int i = 1; i < len; ++i) a[i] -= x[i] * vAj;
            }
        }
        /* B_right[j+1:end,j] -= L_right[j+1:end,j]   :160-162  (fill evaluated with the current U) */
        int nright = jr2_end - jr2_beg + 1; if (nright < 0) nright = 0;
        for (int j = 1; j <= nright; ++j)
            for (int i = k + j; i <= kr_end; ++i) *bptr(A, i, jr2_beg + j - 1) -= fill_get(A, i, jr2_beg + j - 1);
        /* lmul!(Q', B_right)  :163  — banded QRPackedQ' lmul! (BandedMatrices, restated).  Entries of
         * B_right above the band are structurally zero and stay exactly zero, so
         * they are skipped here instead of being written as 0.0. */
        int nrefl = (kr_end - k + 1 < jr1_end - k + 1) ? kr_end - k + 1 : jr1_end - k + 1; /* min(mA,nA) */
        for (int kk = k; kk < k + nrefl; ++kk) {
            int re = (kk + l < kr_end) ? kk + l : kr_end;
            double t = tau[kk - 1];                                /* tau[n] stays 0: no reflector there */
            if (kk > jr3_end) t = 0.0;
            for (int j = jr2_beg; j <= jr2_end; ++j) {
                double vBj = band_get(A, kk, j);
                for (int i = kk + 1; i <= re; ++i) vBj += *bptr(A, i, kk) * band_get(A, i, j);
                vBj = t * vBj;
                if (inband(A, kk, j)) *bptr(A, kk, j) -= vBj;
                for (int i = kk + 1; i <= re; ++i)
                    if (inband(A, i, j)) *bptr(A, i, j) -= *bptr(A, i, kk) * vBj;
            }
        }
        /* U,V = arguments(L_right); lmul!(Q', U)  :164-165  — same reflectors onto rows kr of U */
        for (int kk = k; kk < k + nrefl; ++kk) {
            int re = (kk + l < kr_end) ? kk + l : kr_end;
            double t = (kk > jr3_end) ? 0.0 : tau[kk - 1];
            for (int q = 1; q <= r; ++q) {
                double vBj = Uget(A, kk, q);
                for (int i = kk + 1; i <= re; ++i) vBj += *bptr(A, i, kk) * Uget(A, i, q);
                vBj = t * vBj;
                *Uptr(A, kk, q) -= vBj;
                for (int i = kk + 1; i <= re; ++i) *Uptr(A, i, q) -= *bptr(A, i, kk) * vBj;
            }
        }
        /* B_right[j+1:end,j] += L_right[j+1:end,j]   :166-168  (fill evaluated with the updated U) */
        for (int j = 1; j <= nright; ++j)
            for (int i = k + j; i <= kr_end; ++i) *bptr(A, i, jr2_beg + j - 1) += fill_get(A, i, jr2_beg + j - 1);
        k = jr1_end + 1;                                           /* :169 */
    }
}

/* b <- Q'b with Q = QRPackedQ(bandpart(factors), tau)  (src/AlmostBandedMatrix.jl:181,189; banded lmul! restated) */
static void ab_lmul_qt(const ab_t *F, const double *tau, double *b) {
    int n = F->n, l = F->l;
    for (int k = 1; k <= n; ++k) {
        int re = (k + l < n) ? k + l : n;
        double vb = b[k - 1];
        for (int i = k + 1; i <= re; ++i) vb += *bptr(F, i, k) * b[i - 1];
        vb = tau[k - 1] * vb;
        b[k - 1] -= vb;
        for (int i = k + 1; i <= re; ++i) b[i - 1] -= *bptr(F, i, k) * vb;
    }
}
/* b <- Q b (reverse order, same reflectors); used by the Q*Q' round-trip tests (test_almostbanded.jl:157) */
static void ab_lmul_q(const ab_t *F, const double *tau, double *b) {
    int n = F->n, l = F->l;
    for (int k = n; k >= 1; --k) {
        int re = (k + l < n) ? k + l : n;
        double vb = b[k - 1];
        for (int i = k + 1; i <= re; ++i) vb += *bptr(F, i, k) * b[i - 1];
        vb = tau[k - 1] * vb;
        b[k - 1] -= vb;
        for (int i = k + 1; i <= re; ++i) b[i - 1] -= *bptr(F, i, k) * vb;
    }
}

/* src/AlmostBandedMatrix.jl:215-240.  R = AlmostBanded with bandwidths (l,u).  unit != 0 -> UnitUpperTriangular. */
static void ab_upper_ldiv(const ab_t *R, double *b, int unit) {
    int n = R->n, u = R->u, r = R->r;
    double *buffer = (double *)calloc((size_t)(r > 0 ? r : 1), sizeof(double));        /* :219 */
    int k = n;
    while (k > 0) {
        int kr_beg = (k - u > 1) ? k - u : 1;                       /* kr  = max(1,k-u):k        :225 */
        int jr1_beg = k + 1, jr1_end = k + u + 1;                   /* jr1 = k+1:k+u+1           :226 */
        int jr2_beg = k + u + 2, jr2_end = k + 2 * u + 2;           /* jr2 = k+u+2:k+2u+2        :227 */
        if (jr2_beg < n) {                                          /* :229 (sic: '<') */
            for (int j = jr2_beg; j <= jr2_end && j <= n; ++j)      /* buffer += V[:,jr2]*b[jr2] :230 */
                for (int q = 1; q <= r; ++q) buffer[q - 1] += Vget(R, q, j) * b[j - 1];
            for (int q = 1; q <= r; ++q) {                          /* b[kr] -= U[kr,:]*buffer   :231 */
                double ax = -1.0 * buffer[q - 1];                   /* (ArrayLayouts column-major matvec: */
                for (int i = kr_beg; i <= k; ++i) b[i - 1] += ax * Uget(R, i, q); /* y += (alpha x_j) A[:,j]) */
            }
        }
        if (jr1_beg < n) {                                          /* b[kr] -= R[kr,jr1]*b[jr1] :233-235 */
            for (int j = jr1_beg; j <= jr1_end && j <= n; ++j) {
                double ax = -1.0 * b[j - 1];
                for (int i = kr_beg; i <= k; ++i) b[i - 1] += ax * ab_getindex(R, i, j);
            }
        }
        for (int i = k; i >= kr_beg; --i) {                         /* Tri(B[kr,kr]) \ b[kr]     :236 */
            double s = b[i - 1];
            for (int j = i + 1; j <= k; ++j) s -= band_get(R, i, j) * b[j - 1];
            b[i - 1] = unit ? s : s / *bptr(R, i, i);
        }
        k = kr_beg - 1;                                             /* :237 */
    }
    free(buffer);
}

/* src/AlmostBandedMatrix.jl:258-270 — LowerTriangular / UnitLowerTriangular solves touch only the band part */
static void ab_lower_ldiv(const ab_t *A, double *b, int unit) {
    int n = A->n, l = A->l;
    for (int i = 1; i <= n; ++i) {
        double s = b[i - 1];
        for (int j = (i - l > 1 ? i - l : 1); j < i; ++j) s -= *bptr(A, i, j) * b[j - 1];
        b[i - 1] = unit ? s : s / *bptr(A, i, i);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Exported C entry points (ctypes / bench).  All pointers are plain host arrays in the reference layout.
 * ---------------------------------------------------------------------------------------------- */

/* qr(A): returns widened factors (2l+u+1) x n, updated U (n x r) and tau (n).  Input is not mutated
 * (test_almostbanded.jl:146).  ncols <= 0 selects the reference default min(m-1, n) for real square input (:137). */
int ssm_oracle_ab_qr(int n, int l, int u, int r, const double *bands, const double *U, const double *V,
                     double *factors, double *Uout, double *tau, int ncols) {
    if (n < 1 || l < 0 || u < 0 || r < 0) return -1;
    ab_t A = { n, l, u, r, (double *)bands, (double *)U, (double *)V };
    ab_t W = { n, l, l + u, r, factors, Uout, (double *)V };
    ab_widen(&A, &W);
    memset(tau, 0, sizeof(double) * (size_t)n);
    if (ncols <= 0) ncols = n - 1;
    if (ncols > n) ncols = n;
    ab_qr_inplace(&W, tau, ncols);
    return 0;
}

/* ldiv!(F, b) (src/AlmostBandedMatrix.jl:187-192): b <- Q'b then b <- R \ b.   ub = l + u of the factors. */
int ssm_oracle_ab_ldiv(int n, int l, int ub, int r, const double *factors, const double *Uf, const double *V,
                       const double *tau, double *b) {
    ab_t F = { n, l, ub, r, (double *)factors, (double *)Uf, (double *)V };
    ab_lmul_qt(&F, tau, b);
    ab_upper_ldiv(&F, b, 0);
    return 0;
}
int ssm_oracle_ab_lmul_qt(int n, int l, int ub, const double *factors, const double *tau, double *b) {
    ab_t F = { n, l, ub, 0, (double *)factors, NULL, NULL };
    ab_lmul_qt(&F, tau, b);
    return 0;
}
int ssm_oracle_ab_lmul_q(int n, int l, int ub, const double *factors, const double *tau, double *b) {
    ab_t F = { n, l, ub, 0, (double *)factors, NULL, NULL };
    ab_lmul_q(&F, tau, b);
    return 0;
}
/* Triangular solves on an AlmostBandedMatrix with bandwidths (l,u): uplo 'U'/'L', unit 0/1 (:242-270) */
int ssm_oracle_ab_tri_ldiv(int n, int l, int u, int r, const double *bands, const double *U, const double *V,
                           double *b, int upper, int unit) {
    ab_t A = { n, l, u, r, (double *)bands, (double *)U, (double *)V };
    if (upper) ab_upper_ldiv(&A, b, unit); else ab_lower_ldiv(&A, b, unit);
    return 0;
}

/* A \ b (src/AlmostBandedMatrix.jl:125: _factorize = qr; then ldiv!).  work: (2l+u+1)*n + n*r + n doubles. */
int ssm_oracle_ab_solve(int n, int l, int u, int r, const double *bands, const double *U, const double *V,
                        const double *b, double *x, double *work) {
    double *factors = work, *Uf = factors + (size_t)(2 * l + u + 1) * (size_t)n, *tau = Uf + (size_t)n * (size_t)r;
    int rc = ssm_oracle_ab_qr(n, l, u, r, bands, U, V, factors, Uf, tau, 0);
    if (rc) return rc;
    if (x != b) memcpy(x, b, sizeof(double) * (size_t)n);
    return ssm_oracle_ab_ldiv(n, l, l + u, r, factors, Uf, V, tau, x);
}

/* dense A (n x n column-major) from the structured representation; for the dense LAPACK cross-checks */
int ssm_oracle_ab_dense(int n, int l, int u, int r, const double *bands, const double *U, const double *V, double *out) {
    ab_t A = { n, l, u, r, (double *)bands, (double *)U, (double *)V };
    for (int j = 1; j <= n; ++j)
        for (int k = 1; k <= n; ++k) out[(size_t)(k - 1) + (size_t)n * (size_t)(j - 1)] = ab_getindex(&A, k, j);
    return 0;
}

/* relative residual ||b - A x||_2 / ||b||_2 using getindex semantics in O(n (l+u+r)) */
double ssm_oracle_ab_relres(int n, int l, int u, int r, const double *bands, const double *U, const double *V,
                            const double *x, const double *b) {
    ab_t A = { n, l, u, r, (double *)bands, (double *)U, (double *)V };
    double *s = (double *)calloc((size_t)(r > 0 ? r : 1) * (size_t)(n + 1), sizeof(double)); /* suffix sums of V x */
    for (int j = n; j >= 1; --j)
        for (int q = 1; q <= r; ++q) s[(size_t)(q - 1) + (size_t)r * (size_t)(j - 1)] = s[(size_t)(q - 1) + (size_t)r * (size_t)j] + Vget(&A, q, j) * x[j - 1];
    double num = 0.0, den = 0.0;
    for (int k = 1; k <= n; ++k) {
        double ax = 0.0;
        for (int j = (k - l > 1 ? k - l : 1); j <= (k + u < n ? k + u : n); ++j) ax += *bptr(&A, k, j) * x[j - 1];
        if (k + u + 1 <= n)
            for (int q = 1; q <= r; ++q) ax += Uget(&A, k, q) * s[(size_t)(q - 1) + (size_t)r * (size_t)(k + u)];
        double d = b[k - 1] - ax;
        num += d * d; den += b[k - 1] * b[k - 1];
    }
    free(s);
    return sqrt(num) / sqrt(den);
}

/* ---- synthetic inputs in the reference layout (same generator as the device kernel) ---- */
int ssm_oracle_ab_generate(int n, int l, int u, int r, int64_t s0, int64_t count, uint64_t seed, int dist,
                           double *bands, double *U, double *V, double *b) {
    size_t bs = (size_t)(l + u + 1) * (size_t)n, us = (size_t)n * (size_t)r;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < count; ++c) {
        uint64_t key = ssm_syskey(seed, (uint64_t)(s0 + c));
        if (bands) for (int j = 0; j < n; ++j) for (int d = 0; d <= l + u; ++d)
            bands[bs * (size_t)c + (size_t)d + (size_t)(l + u + 1) * (size_t)j] = ssm_ab_band_elem(key, dist, n, l, u, r, d, j);
        if (U) for (int q = 0; q < r; ++q) for (int i = 0; i < n; ++i)
            U[us * (size_t)c + (size_t)i + (size_t)n * (size_t)q] = ssm_ab_U_elem(key, dist, n, r, i, q);
        if (V) for (int j = 0; j < n; ++j) for (int q = 0; q < r; ++q)
            V[us * (size_t)c + (size_t)q + (size_t)r * (size_t)j] = ssm_ab_V_elem(key, dist, n, r, q, j);
        if (b) for (int i = 0; i < n; ++i) b[(size_t)n * (size_t)c + (size_t)i] = ssm_ab_b_elem(key, i);
    }
    return 0;
}

/* ---- batched drivers (OpenMP over independent systems): the CPU baseline of SURVEY 8(d) ---- */
int ssm_oracle_ab_solve_batch(int n, int l, int u, int r, int64_t nb, const double *bands, const double *U,
                              const double *V, const double *b, double *x) {
    size_t bs = (size_t)(l + u + 1) * (size_t)n, us = (size_t)n * (size_t)r;
    size_t wsz = (size_t)(2 * l + u + 1) * (size_t)n + us + (size_t)n;
    int rc = 0;
#pragma omp parallel
    {
        double *work = (double *)malloc(sizeof(double) * wsz);
#pragma omp for schedule(static)
        for (int64_t s = 0; s < nb; ++s) {
            int e = ssm_oracle_ab_solve(n, l, u, r, bands + bs * (size_t)s, U + us * (size_t)s, V + us * (size_t)s,
                                        b + (size_t)n * (size_t)s, x + (size_t)n * (size_t)s, work);
            if (e) rc = e;
        }
        free(work);
    }
    return rc;
}
/* qr + ldiv! as two separately stored operations (the headline "QR+solve" accounting): factors are kept */
int ssm_oracle_ab_qr_batch(int n, int l, int u, int r, int64_t nb, const double *bands, const double *U,
                           const double *V, double *factors, double *Uout, double *tau) {
    size_t bs = (size_t)(l + u + 1) * (size_t)n, us = (size_t)n * (size_t)r, fs = (size_t)(2 * l + u + 1) * (size_t)n;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < nb; ++s)
        ssm_oracle_ab_qr(n, l, u, r, bands + bs * (size_t)s, U + us * (size_t)s, V + us * (size_t)s,
                         factors + fs * (size_t)s, Uout + us * (size_t)s, tau + (size_t)n * (size_t)s, 0);
    return 0;
}
int ssm_oracle_ab_ldiv_batch(int n, int l, int ub, int r, int64_t nb, const double *factors, const double *Uf,
                             const double *V, const double *tau, double *b) {
    size_t us = (size_t)n * (size_t)r, fs = (size_t)(l + ub + 1) * (size_t)n;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < nb; ++s)
        ssm_oracle_ab_ldiv(n, l, ub, r, factors + fs * (size_t)s, Uf + us * (size_t)s, V + us * (size_t)s,
                           tau + (size_t)n * (size_t)s, b + (size_t)n * (size_t)s);
    return 0;
}

/* ---- OpenMP team size of the batched drivers (bench.py's cpu_baseline / --impl reference legs) ----
 * want > 0: omp_set_num_threads(want) first — torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which would
 * silently turn the "all host cores" baseline into a single-thread one.  Returns the team size a parallel region really gets. */
#ifdef _OPENMP
#include <omp.h>
#endif
int ssm_oracle_threads(int want) {
    int got = 1;
#ifdef _OPENMP
    if (want > 0) { omp_set_dynamic(0); omp_set_num_threads(want); }
#pragma omp parallel
    {
#pragma omp single
        got = omp_get_num_threads();
    }
#else
    (void)want;
#endif
    return got;
}
