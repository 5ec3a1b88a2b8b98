/* bps_oracle.c — CPU restatement (plain C) of the reference's BandedPlusSemiseparable QR sweep and solves.
 *
 * TEST INFRASTRUCTURE ONLY (see ab_oracle.c).  Restates, term by term and in the reference's own order:
 *   bps_getindex0                src/BandedPlusSemiseparableMatrix.jl:43-51   A = B + tril(UV',-1) + triu(WS',1)
 *   bps_mul                      src/BandedPlusSemiseparableMatrix.jl:21-41   O(n) matvec
 *   bps_upper_ldiv               src/BandedPlusSemiseparableMatrix.jl:53-70   ldiv!(UpperTriangular{BPS}, b)
 *   fast_UtA                     src/semiseparableqr.jl:620-640
 *   state constructor            src/semiseparableqr.jl:36-48                 B->(l,l+m), W->[W 0], S->[S A'U]
 *   lookup tables                src/semiseparableqr.jl:573-618               UtU, uw_sum, d_extra
 *   compute_Householder_vector   src/semiseparableqr.jl:315-364
 *   compute_d1_and_dbar          src/semiseparableqr.jl:366-393
 *   compute_variables_c          src/semiseparableqr.jl:395-438
 *   update_next_submatrix!       src/semiseparableqr.jl:440-512
 *   update_upper_triangular_part! src/semiseparableqr.jl:514-562
 *   update_lower_triangular_part! src/semiseparableqr.jl:564-571
 *   bandedplussemi_qr! / qr      src/semiseparableqr.jl:104-131 (incl. the final A.B[n,n] = A[n,n], :123, via getindex :50-102)
 *   lmul!(Q', b)                 src/semiseparableqr.jl:863-919
 *
 * Parity pinning: the reference pins this path against LinearAlgebra.qrfactUnblocked!(Matrix(A)) (LAPACK geqr2
 * convention) and explicit dense products (test/test_qr.jl:13-53, test/test_bandedplussemiseparable.jl:11-17) on
 * unseeded randn data; tests/test_oracle_bps.py re-instantiates those checks with seeded data against
 * scipy's dgeqrf.  Julia is not available in this image.
 *
 * Layout = the reference's: B.data (l+m+1) x n column-major band storage (bandwidths (l,m)); U,V n x r and
 * W,S n x p column-major.  Indices of the static helpers are 1-based.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/ssm_rng.h"

#define MAXR 8
#define MAXB 40   /* l+m+1 and friends */

typedef struct {
    int n, l, mw;      /* B has bandwidths (l, mw); mw = m on input, l+m in the factor state */
    int r, pw;         /* U,V: n x r;  W,S: n x pw (pw = p on input, p+r in the factor state) */
    double *B, *U, *V, *W, *S;
} bps_t;

static inline int b_in(const bps_t *A, int k, int j) { return (j - k <= A->mw) && (k - j <= A->l) && k >= 1 && j >= 1 && k <= A->n && j <= A->n; }
static inline double *Bp(const bps_t *A, int k, int j) { return &A->B[(size_t)(A->mw + k - j) + (size_t)(A->l + A->mw + 1) * (size_t)(j - 1)]; }
static inline double Bg(const bps_t *A, int k, int j) { return b_in(A, k, j) ? *Bp(A, k, j) : 0.0; }
#define M2(P, ld, i, j) ((P)[(size_t)((i) - 1) + (size_t)(ld) * (size_t)((j) - 1)])
#define Ug(A, i, q) M2((A)->U, (A)->n, i, q)
#define Vg(A, i, q) M2((A)->V, (A)->n, i, q)
#define Wg(A, i, q) M2((A)->W, (A)->n, i, q)
#define Sg(A, i, q) M2((A)->S, (A)->n, i, q)
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* src/BandedPlusSemiseparableMatrix.jl:43-51 */
static double bps_getindex0(const bps_t *A, int k, int j) {
    double s = 0.0;
    if (j > k) { for (int q = 1; q <= A->pw; ++q) s += Wg(A, k, q) * Sg(A, j, q); return s + Bg(A, k, j); }
    if (k > j) { for (int q = 1; q <= A->r; ++q) s += Ug(A, k, q) * Vg(A, j, q); return s + Bg(A, k, j); }
    return Bg(A, k, j);
}

/* src/BandedPlusSemiseparableMatrix.jl:21-41 */
static void bps_mul(const bps_t *A, const double *b, double *res) {
    int n = A->n, r = A->r, p = A->pw, l = A->l, m = A->mw;
    double Stb[MAXR], Vtb[MAXR];
    for (int q = 1; q <= p; ++q) { double s = 0.0; for (int i = 1; i <= n; ++i) s += Sg(A, i, q) * b[i - 1]; Stb[q - 1] = s; }
    for (int q = 0; q < r; ++q) Vtb[q] = 0.0;
    for (int k = 1; k <= n; ++k) {
        double Bb = 0.0;
        for (int j = imax(1, k - l); j <= imin(n, k + m); ++j) Bb += *Bp(A, k, j) * b[j - 1];
        for (int q = 1; q <= p; ++q) Stb[q - 1] += -b[k - 1] * Sg(A, k, q);
        double uv = 0.0, ws = 0.0;
        for (int q = 1; q <= r; ++q) uv += Ug(A, k, q) * Vtb[q - 1];
        for (int q = 1; q <= p; ++q) ws += Wg(A, k, q) * Stb[q - 1];
        res[k - 1] = uv + Bb + ws;
        for (int q = 1; q <= r; ++q) Vtb[q - 1] += b[k - 1] * Vg(A, k, q);
    }
}

/* src/BandedPlusSemiseparableMatrix.jl:53-70 */
static void bps_upper_ldiv(const bps_t *F, double *b) {
    int n = F->n, p = F->pw, m = F->mw;
    double sx[2 * MAXR];
    for (int q = 0; q < p; ++q) sx[q] = 0.0;
    for (int j = n; j >= 1; --j) {
        double residual = 0.0;
        for (int q = 1; q <= p; ++q) residual += Wg(F, j, q) * sx[q - 1];
        for (int k = j + 1; k <= imin(j + m, n); ++k) residual += *Bp(F, j, k) * b[k - 1];
        b[j - 1] = (b[j - 1] - residual) / *Bp(F, j, j);
        for (int q = 1; q <= p; ++q) sx[q - 1] += b[j - 1] * Sg(F, j, q);
    }
}

/* src/semiseparableqr.jl:620-640.  out: r x (n+1-j) column-major.  p = columns of W/S used (= A->pw) */
static void fast_UtA(const bps_t *A, int j, double *out) {
    int n = A->n, r = A->r, p = A->pw, l = A->l, m = A->mw;
    double UtU[MAXR * MAXR], UtW[MAXR * 2 * MAXR];
    for (int a = 1; a <= r; ++a) for (int b = 1; b <= r; ++b) { double s = 0.0; for (int i = j; i <= n; ++i) s += Ug(A, i, a) * Ug(A, i, b); M2(UtU, r, a, b) = s; }
    for (int a = 0; a < r * p; ++a) UtW[a] = 0.0;
    for (int i = j; i <= n; ++i) {
        for (int b = 1; b <= r; ++b) for (int a = 1; a <= r; ++a) M2(UtU, r, a, b) += -1.0 * Ug(A, i, a) * Ug(A, i, b);   /* :631 */
        double *col = out + (size_t)r * (size_t)(i - j);
        for (int a = 1; a <= r; ++a) { double s = 0.0; for (int b = 1; b <= r; ++b) s += M2(UtU, r, a, b) * Vg(A, i, b); col[a - 1] = s; }      /* :633 */
        int lo = imax(j, i - m), hi = imin(i + l, n);
        for (int a = 1; a <= r; ++a) { double s = 0.0; for (int t = lo; t <= hi; ++t) s += Ug(A, t, a) * *Bp(A, t, i); col[a - 1] += s; }       /* :634 */
        for (int a = 1; a <= r; ++a) { double s = 0.0; for (int q = 1; q <= p; ++q) s += M2(UtW, r, a, q) * Sg(A, i, q); col[a - 1] += s; }     /* :635 */
        for (int q = 1; q <= p; ++q) for (int a = 1; a <= r; ++a) M2(UtW, r, a, q) += Ug(A, i, a) * Wg(A, i, q);                               /* :637 */
    }
}

/* ---------------------------------------------------------------------------------------------------------
 * qr(A::BandedPlusSemiseparableMatrix)  (src/semiseparableqr.jl:128-131)
 * Inputs (not mutated): B (l+m+1) x n, U,V n x r, W,S n x p.
 * Outputs: Bf (2l+m+1) x n band storage with bandwidths (l, l+m); Vf n x r; Wf, Sf n x (p+r); tau n.  U unchanged.
 * ------------------------------------------------------------------------------------------------------- */
static int bps_qr_core(int n, int l, int m, int r, int p, const double *B, const double *U, const double *V, const double *W,
                       const double *S, double *Bf, double *Vf, double *Wf, double *Sf, double *tau, int nsteps, double *state);
int ssm_oracle_bps_qr(int n, int l, int m, int r, int p, const double *B, const double *U, const double *V, const double *W,
                      const double *S, double *Bf, double *Vf, double *Wf, double *Sf, double *tau) {
    return bps_qr_core(n, l, m, r, p, B, U, V, W, S, Bf, Vf, Wf, Sf, tau, n - 1, NULL);
}
/* The sweep stopped after `nsteps` columns, with the state of BandedPlusSemiseparableQRPerturbedFactors (src/semiseparableqr.jl:14-48) written to
 * `state` = Q (r x p) | K (r x r) | Es (r x lm) | Xs (lx x p) | Ys (lx x r) | Zs (lx x lm), each column-major, lm = min(l+m, n), lx = min(l, n).
 * For tests of the invariant the reference states at :1-12 and checks in test/test_getindex.jl:32-38: after j steps the trailing block is
 *   A0 + U Q Sp' + U K U'A0 + U [Es 0] + [Xs;0] Sp' + [Ys;0] U'A0 + [Zs 0;0 0]. */
int ssm_oracle_bps_qr_steps(int n, int l, int m, int r, int p, const double *B, const double *U, const double *V, const double *W,
                            const double *S, int nsteps, double *Bf, double *Vf, double *Wf, double *Sf, double *tau, double *state) {
    if (nsteps < 0 || nsteps > n - 1 || !state) return -1;
    return bps_qr_core(n, l, m, r, p, B, U, V, W, S, Bf, Vf, Wf, Sf, tau, nsteps, state);
}
static int bps_qr_core(int n, int l, int m, int r, int p, const double *B, const double *U, const double *V, const double *W,
                       const double *S, double *Bf, double *Vf, double *Wf, double *Sf, double *tau, int nsteps, double *state) {
    if (n < 1 || l < 0 || m < 0 || r < 0 || p < 0 || r > MAXR || p > MAXR || 2 * l + m + 1 > MAXB) return -1;
    const int pr = p + r, lm = imin(l + m, n), lx = imin(l, n);
    /* ---- constructor :36-48 ---- */
    bps_t A0 = { n, l, m, r, p, (double *)B, (double *)U, (double *)V, (double *)W, (double *)S };
    double *AtU = (double *)malloc(sizeof(double) * (size_t)r * (size_t)n + 8);
    fast_UtA(&A0, 1, AtU);                                                     /* r x n; AtU' is n x r        :42 */
    bps_t A = { n, l, l + m, r, pr, Bf, (double *)U, Vf, Wf, Sf };
    memset(Bf, 0, sizeof(double) * (size_t)(2 * l + m + 1) * (size_t)n);
    for (int j = 1; j <= n; ++j) for (int k = imax(1, j - m); k <= imin(n, j + l); ++k) *Bp(&A, k, j) = *Bp(&A0, k, j);   /* BandedMatrix(B,(l,l+m)) */
    memcpy(Vf, V, sizeof(double) * (size_t)n * (size_t)r);
    for (int q = 1; q <= p; ++q) for (int i = 1; i <= n; ++i) { M2(Wf, n, i, q) = M2(W, n, i, q); M2(Sf, n, i, q) = M2(S, n, i, q); }
    for (int q = 1; q <= r; ++q) for (int i = 1; i <= n; ++i) { M2(Wf, n, i, p + q) = 0.0; M2(Sf, n, i, p + q) = M2(AtU, r, q, i); }     /* [W 0], [S A'U] :43 */
    double Q[MAXR * MAXR] = {0}, K[MAXR * MAXR] = {0}, E[MAXR * MAXB] = {0}, X[MAXB * MAXR] = {0}, Y[MAXB * MAXR] = {0}, Z[MAXB * MAXB] = {0};
    double Qp[MAXR * MAXR], Kp[MAXR * MAXR], Ep[MAXR * MAXB], Xp[MAXB * MAXR], Yp[MAXB * MAXR], Zp[MAXB * MAXB];
#define Qm(P, a, b) M2(P, r, a, b)   /* r x p  */
#define Km(P, a, b) M2(P, r, a, b)   /* r x r  */
#define Em(P, a, t) M2(P, r, a, t)   /* r x lm */
#define Xm(P, s, b) M2(P, lx, s, b)  /* lx x p */
#define Ym(P, s, b) M2(P, lx, s, b)  /* lx x r */
#define Zm(P, s, t) M2(P, lx, s, t)  /* lx x lm */
    memset(tau, 0, sizeof(double) * (size_t)n);
    /* ---- lookup tables :573-618 ---- */
    double *UtU = (double *)calloc((size_t)(n + 2) * r * r + 1, sizeof(double));      /* UtU[i] = sum_{t>=i} u_t u_t' */
    double *uw = (double *)calloc((size_t)(n + 2) * r * p + 1, sizeof(double));        /* uw[t]  = sum_{s<t} u_s w_s'   */
    double *dex = (double *)calloc((size_t)(n + 2) * r * (m > 0 ? m : 1) + 1, sizeof(double));
#define UTU(i) (UtU + (size_t)(i) * r * r)
#define UW(i) (uw + (size_t)(i) * r * p)
#define DEX(i) (dex + (size_t)(i) * r * (m > 0 ? m : 1))
    {
        double cur[MAXR * MAXR] = {0};
        for (int i = n; i >= 1; --i) {
            for (int b = 1; b <= r; ++b) for (int a = 1; a <= r; ++a) M2(cur, r, a, b) += Ug(&A, i, a) * Ug(&A, i, b);
            memcpy(UTU(i), cur, sizeof(double) * r * r);
        }
        double cw[MAXR * MAXR] = {0};
        for (int t = 1; t <= n; ++t) {
            memcpy(UW(t), cw, sizeof(double) * r * p);
            for (int q = 1; q <= p; ++q) for (int a = 1; a <= r; ++a) M2(cw, r, a, q) += Ug(&A, t, a) * Wg(&A, t, q);
        }
        for (int i = 1; i <= n; ++i) {
            int nc = imin(m, n - i);
            for (int t = imax(1, i + 1 - m); t <= i - 1; ++t)
                for (int c = 1; c <= nc; ++c) for (int a = 1; a <= r; ++a) M2(DEX(i), r, a, c) += Ug(&A, t, a) * Bg(&A, t, i + c);
        }
    }
    /* ---- n-1 Householder steps :117-121, onestep_qr! :133-152 ---- */
    for (int j = 0; j < nsteps; ++j) {
        const int q1 = j + 1;                         /* current column / row */
        const double *G = UTU(j + 2);
        /* == compute_Householder_vector :315-364 == */
        double UtA1[MAXR], kb[MAXR], b[MAXB];
        int rlo = j + 1, rhi = imin(l + j + 1, n);
        for (int a = 1; a <= r; ++a) { double s = 0.0; for (int t = rlo; t <= rhi; ++t) s += Ug(&A, t, a) * *Bp(&A, t, q1); UtA1[a - 1] = s; }         /* :326 */
        for (int a = 1; a <= r; ++a) { double s = 0.0; for (int c = 1; c <= r; ++c) s += M2(G, r, a, c) * Vg(&A, q1, c); UtA1[a - 1] += s; }           /* :327 */
        for (int a = 1; a <= r; ++a) { double s = 0.0; for (int c = 1; c <= r; ++c) s += Km(K, a, c) * UtA1[c - 1]; kb[a - 1] = s; }                  /* :330 */
        for (int a = 1; a <= r; ++a) { double s = 0.0; for (int c = 1; c <= p; ++c) s += Qm(Q, a, c) * Sg(&A, q1, c); kb[a - 1] += s; }               /* :331 */
        for (int a = 1; a <= r; ++a) kb[a - 1] += Vg(&A, q1, a);                                                                                      /* :332 */
        if (l > 0 || m > 0) for (int a = 1; a <= r; ++a) kb[a - 1] += Em(E, a, 1);                                                                     /* :333-335 */
        const int lb = imin(l + j + 1, n) - (j + 2) + 1 > 0 ? imin(l + j + 1, n) - (j + 2) + 1 : 0;
        for (int s = 1; s <= lb; ++s) b[s - 1] = *Bp(&A, j + 1 + s, q1);                                                                             /* :336 */
        if (l > 0 || m > 0) {
            int krhi = imin(l, lb + 1);
            for (int kr = 2; kr <= krhi; ++kr) {                                                                                                    /* :339-343 */
                double s1 = 0.0, s2 = 0.0;
                for (int c = 1; c <= p; ++c) s1 += Xm(X, kr, c) * Sg(&A, q1, c);
                for (int c = 1; c <= r; ++c) s2 += Ym(Y, kr, c) * UtA1[c - 1];
                b[kr - 2] += s1; b[kr - 2] += s2; b[kr - 2] += Zm(Z, kr, 1);
            }
        }
        double pivot;
        {
            double t1[MAXR], s1 = 0.0, s2 = 0.0;   /* (u1' Q) S1p  and  (u1' K) UtA1 :346 */
            for (int c = 1; c <= p; ++c) { double s = 0.0; for (int a = 1; a <= r; ++a) s += Ug(&A, q1, a) * Qm(Q, a, c); t1[c - 1] = s; }
            for (int c = 1; c <= p; ++c) s1 += t1[c - 1] * Sg(&A, q1, c);
            for (int c = 1; c <= r; ++c) { double s = 0.0; for (int a = 1; a <= r; ++a) s += Ug(&A, q1, a) * Km(K, a, c); t1[c - 1] = s; }
            for (int c = 1; c <= r; ++c) s2 += t1[c - 1] * UtA1[c - 1];
            pivot = *Bp(&A, q1, q1) + s1 + s2;
            if (l > 0) {                                                                                                                        /* :347-348 */
                double e = 0.0, x = 0.0, y = 0.0;
                for (int a = 1; a <= r; ++a) e += Ug(&A, q1, a) * Em(E, a, 1);
                for (int c = 1; c <= p; ++c) x += Xm(X, 1, c) * Sg(&A, q1, c);
                for (int c = 1; c <= r; ++c) y += Ym(Y, 1, c) * UtA1[c - 1];
                pivot += e + x + y + Zm(Z, 1, 1);
            } else if (m > 0) {
                double e = 0.0;
                for (int a = 1; a <= r; ++a) e += Ug(&A, q1, a) * Em(E, a, 1);
                pivot += e;
            }
        }
        double len_square;
        {
            double t1[MAXR], s = 0.0;                                                                                                              /* :354 */
            for (int c = 1; c <= r; ++c) { double z = 0.0; for (int a = 1; a <= r; ++a) z += kb[a - 1] * M2(G, r, a, c); t1[c - 1] = z; }
            for (int c = 1; c <= r; ++c) s += t1[c - 1] * kb[c - 1];
            len_square = pivot * pivot + s;
            if (l > 0) {                                                                                                                        /* :355-357 */
                double kUb = 0.0, bUk = 0.0, bb = 0.0, tk[MAXR];
                for (int t = 1; t <= lb; ++t) { double z = 0.0; for (int a = 1; a <= r; ++a) z += kb[a - 1] * Ug(&A, j + 1 + t, a); kUb += z * b[t - 1]; }
                for (int a = 1; a <= r; ++a) { double z = 0.0; for (int t = 1; t <= lb; ++t) z += b[t - 1] * Ug(&A, j + 1 + t, a); tk[a - 1] = z; }
                for (int a = 1; a <= r; ++a) bUk += tk[a - 1] * kb[a - 1];
                for (int t = 1; t <= lb; ++t) bb += b[t - 1] * b[t - 1];
                len_square += kUb + bUk + bb;
            }
        }
        const double sgn = (pivot > 0) - (pivot < 0);
        const double pivot_new = -sgn * sqrt(len_square);                                                                                       /* :359 */
        const double den = pivot - pivot_new;
        for (int a = 0; a < r; ++a) kb[a] /= den;                                                                                                 /* :360 */
        for (int t = 0; t < lb; ++t) b[t] /= den;                                                                                                 /* :361 */
        const double tc = 2 * (den * den) / (den * den + (len_square - pivot * pivot));                                                          /* :362 */
        tau[j] = tc;                                                                                                                              /* :135 */
        /* == compute_d1_and_dbar :366-393 == */
        double w1[MAXR], u1[MAXR], d1[MAXB], f[MAXR], dbar[MAXB];
        for (int c = 1; c <= p; ++c) w1[c - 1] = Wg(&A, q1, c);
        for (int a = 1; a <= r; ++a) u1[a - 1] = Ug(&A, q1, a);
        const int ld1 = imin(j + 1 + m, n) - (j + 1) + 1;
        for (int t = 1; t <= ld1; ++t) d1[t - 1] = Bg(&A, q1, j + t);                                                                             /* :375 */
        { double s = 0.0; for (int c = 1; c <= p; ++c) s += w1[c - 1] * Sg(&A, q1, c); d1[0] = d1[0] - s; }                                        /* :376 */
        for (int c = 1; c <= p; ++c) { double s = 0.0; for (int t = 1; t <= lb; ++t) s += Wg(&A, j + 1 + t, c) * b[t - 1]; f[c - 1] = s; }          /* :377 */
        const int i2lo = j + 1, i2hi = imin(j + 1 + m + l, n), ldb = i2hi - i2lo + 1;
        for (int tj = 1; tj <= ldb; ++tj) {                                                                                                     /* :380-391 */
            int jj = i2lo + tj - 1;
            double acc = 0.0;
            for (int ti = 1; ti <= lb; ++ti) {
                int ii = j + 1 + ti;
                double a0 = 0.0;
                if (ii - jj >= 1) { for (int a = 1; a <= r; ++a) a0 += Ug(&A, ii, a) * Vg(&A, jj, a); }
                a0 += Bg(&A, ii, jj);
                if (jj - ii >= 1) { double up = 0.0; for (int c = 1; c <= p; ++c) up += Wg(&A, ii, c) * Sg(&A, jj, c); a0 += up; }
                acc += a0 * b[ti - 1];
            }
            double sf = 0.0;
            for (int c = 1; c <= p; ++c) sf += Sg(&A, jj, c) * f[c - 1];
            dbar[tj - 1] = acc + -1.0 * sf;
        }
        /* == compute_variables_c :395-438 == */
        double c1[MAXR], c2[MAXR], c3[MAXR], c4[MAXR], c5[MAXR], c6[MAXB];
        {
            double QG[MAXR * MAXR], QUb[MAXR * MAXB];
            /* c1 = Q'u1 + (Q'G) kb + (Q'Ub') b  :404-406 */
            for (int c = 1; c <= p; ++c) { double s = 0.0; for (int a = 1; a <= r; ++a) s += Qm(Q, a, c) * u1[a - 1]; c1[c - 1] = s; }
            for (int c = 1; c <= p; ++c) for (int d = 1; d <= r; ++d) { double s = 0.0; for (int a = 1; a <= r; ++a) s += Qm(Q, a, c) * M2(G, r, a, d); M2(QG, p, c, d) = s; }
            for (int c = 1; c <= p; ++c) { double s = 0.0; for (int d = 1; d <= r; ++d) s += M2(QG, p, c, d) * kb[d - 1]; c1[c - 1] += s; }
            for (int c = 1; c <= p; ++c) for (int t = 1; t <= lb; ++t) { double s = 0.0; for (int a = 1; a <= r; ++a) s += Qm(Q, a, c) * Ug(&A, j + 1 + t, a); M2(QUb, p, c, t) = s; }
            for (int c = 1; c <= p; ++c) { double s = 0.0; for (int t = 1; t <= lb; ++t) s += M2(QUb, p, c, t) * b[t - 1]; c1[c - 1] += s; }
            /* c2 same with K :408-410 */
            for (int c = 1; c <= r; ++c) { double s = 0.0; for (int a = 1; a <= r; ++a) s += Km(K, a, c) * u1[a - 1]; c2[c - 1] = s; }
            for (int c = 1; c <= r; ++c) for (int d = 1; d <= r; ++d) { double s = 0.0; for (int a = 1; a <= r; ++a) s += Km(K, a, c) * M2(G, r, a, d); M2(QG, r, c, d) = s; }
            for (int c = 1; c <= r; ++c) { double s = 0.0; for (int d = 1; d <= r; ++d) s += M2(QG, r, c, d) * kb[d - 1]; c2[c - 1] += s; }
            for (int c = 1; c <= r; ++c) for (int t = 1; t <= lb; ++t) { double s = 0.0; for (int a = 1; a <= r; ++a) s += Km(K, a, c) * Ug(&A, j + 1 + t, a); M2(QUb, r, c, t) = s; }
            for (int c = 1; c <= r; ++c) { double s = 0.0; for (int t = 1; t <= lb; ++t) s += M2(QUb, r, c, t) * b[t - 1]; c2[c - 1] += s; }
            /* c3 = G kb + u1 + Ub' b :412-414 */
            for (int a = 1; a <= r; ++a) { double s = 0.0; for (int d = 1; d <= r; ++d) s += M2(G, r, a, d) * kb[d - 1]; c3[a - 1] = s; }
            for (int a = 1; a <= r; ++a) c3[a - 1] += u1[a - 1];
            for (int a = 1; a <= r; ++a) { double s = 0.0; for (int t = 1; t <= lb; ++t) s += Ug(&A, j + 1 + t, a) * b[t - 1]; c3[a - 1] += s; }
        }
        const int nv = imin(l, n - j) - 1 > 0 ? imin(l, n - j) - 1 : 0;     /* rows 2..min(l,n-j) of Xs,Ys,Zs  :416-418 */
        const int lc6 = imin(l + m, n - j);
        {
            double T[MAXB * MAXR];
            /* c4 = Xv' b[1:nv] + (Xv' U[j+2:j+1+nv,:]) kb  :420-421 */
            for (int c = 1; c <= p; ++c) { double s = 0.0; for (int t = 1; t <= nv; ++t) s += Xm(X, t + 1, c) * b[t - 1]; c4[c - 1] = s; }
            for (int c = 1; c <= p; ++c) for (int a = 1; a <= r; ++a) { double s = 0.0; for (int t = 1; t <= nv; ++t) s += Xm(X, t + 1, c) * Ug(&A, j + 1 + t, a); M2(T, p, c, a) = s; }
            for (int c = 1; c <= p; ++c) { double s = 0.0; for (int a = 1; a <= r; ++a) s += M2(T, p, c, a) * kb[a - 1]; c4[c - 1] += s; }
            for (int c = 1; c <= r; ++c) { double s = 0.0; for (int t = 1; t <= nv; ++t) s += Ym(Y, t + 1, c) * b[t - 1]; c5[c - 1] = s; }
            for (int c = 1; c <= r; ++c) for (int a = 1; a <= r; ++a) { double s = 0.0; for (int t = 1; t <= nv; ++t) s += Ym(Y, t + 1, c) * Ug(&A, j + 1 + t, a); M2(T, r, c, a) = s; }
            for (int c = 1; c <= r; ++c) { double s = 0.0; for (int a = 1; a <= r; ++a) s += M2(T, r, c, a) * kb[a - 1]; c5[c - 1] += s; }
            for (int c = 1; c <= lc6; ++c) { double s = 0.0; for (int t = 1; t <= nv; ++t) s += Zm(Z, t + 1, c) * b[t - 1]; c6[c - 1] = s; }
            for (int c = 1; c <= lc6; ++c) for (int a = 1; a <= r; ++a) { double s = 0.0; for (int t = 1; t <= nv; ++t) s += Zm(Z, t + 1, c) * Ug(&A, j + 1 + t, a); M2(T, lc6, c, a) = s; }
            for (int c = 1; c <= lc6; ++c) { double s = 0.0; for (int a = 1; a <= r; ++a) s += M2(T, lc6, c, a) * kb[a - 1]; c6[c - 1] += s; }
            if (l > 0) {                                                                                                                        /* :428-435 */
                for (int c = 1; c <= p; ++c) c4[c - 1] += Xm(X, 1, c);
                for (int c = 1; c <= r; ++c) c5[c - 1] += Ym(Y, 1, c);
                for (int c = 1; c <= lc6; ++c) c6[c - 1] += Zm(Z, 1, c);
            }
        }
        /* == copy state :139-144 == */
        memcpy(Qp, Q, sizeof Q); memcpy(Kp, K, sizeof K); memcpy(Ep, E, sizeof E); memcpy(Xp, X, sizeof X); memcpy(Yp, Y, sizeof Y); memcpy(Zp, Z, sizeof Z);
        double c2u = 0.0, c5u = 0.0, ku = 0.0;
        for (int a = 0; a < r; ++a) { c2u += c2[a] * u1[a]; c5u += c5[a] * u1[a]; ku += kb[a] * u1[a]; }
        double Ku1[MAXR];
        for (int a = 1; a <= r; ++a) { double s = 0.0; for (int c = 1; c <= r; ++c) s += Km(Kp, a, c) * u1[c - 1]; Ku1[a - 1] = s; }
        /* == update_next_submatrix! :440-512 == */
        for (int c = 1; c <= p; ++c) for (int a = 1; a <= r; ++a) {
            double v = Qm(Q, a, c);
            v += -tc * kb[a - 1] * w1[c - 1];                                   /* :444 */
            v += -tc * kb[a - 1] * f[c - 1];                                    /* :445 */
            v += -tc * kb[a - 1] * c1[c - 1];                                   /* :446 */
            { double s = 0.0; for (int d = 1; d <= r; ++d) s += Km(Kp, a, d) * (u1[d - 1] * w1[c - 1]); v += s; }   /* :447 */
            v += (-tc * (c2u + c5u)) * kb[a - 1] * w1[c - 1];                   /* :448 */
            v += -tc * kb[a - 1] * c4[c - 1];                                   /* :449 */
            Qm(Q, a, c) = v;
        }
        for (int c = 1; c <= r; ++c) for (int a = 1; a <= r; ++a) {
            double v = Km(K, a, c);
            v += -tc * kb[a - 1] * kb[c - 1]; v += -tc * kb[a - 1] * c2[c - 1]; v += -tc * kb[a - 1] * c5[c - 1];      /* :452-454 */
            Km(K, a, c) = v;
        }
        for (int t = 1; t <= ldb - 1; ++t) for (int a = 1; a <= r; ++a) Em(E, a, t) = -tc * kb[a - 1] * dbar[t];       /* :457 */
        for (int t = 1; t <= ld1 - 1; ++t) for (int a = 1; a <= r; ++a) {
            double v = Em(E, a, t);
            v += -tc * kb[a - 1] * d1[t];                                       /* :459 */
            v += Ku1[a - 1] * d1[t];                                            /* :460 */
            v += -tc * (kb[a - 1] * c2u) * d1[t];                               /* :461 */
            v += -tc * (kb[a - 1] * c5u) * d1[t];                               /* :462 */
            Em(E, a, t) = v;
        }
        for (int t = 1; t <= lm - 1; ++t) for (int a = 1; a <= r; ++a) {
            double v = Em(E, a, t);
            v += Em(Ep, a, t + 1);                                              /* :464 */
            { double s = 0.0; for (int d = 1; d <= r; ++d) s += (kb[a - 1] * c3[d - 1]) * Em(Ep, d, t + 1); v += -tc * s; }   /* :465 */
            Em(E, a, t) = v;
        }
        for (int t = 1; t <= lc6 - 1; ++t) for (int a = 1; a <= r; ++a) Em(E, a, t) += -tc * kb[a - 1] * c6[t];         /* :467 */
        for (int t = ldb; t <= lm; ++t) for (int a = 1; a <= r; ++a) Em(E, a, t) = 0.0;                                 /* :469 */

        for (int c = 1; c <= p; ++c) for (int s = 1; s <= lb; ++s) {
            double v = -tc * b[s - 1] * w1[c - 1];                              /* :472 */
            v += -tc * b[s - 1] * f[c - 1]; v += -tc * b[s - 1] * c1[c - 1];    /* :473-474 */
            v += -tc * b[s - 1] * (c2u * w1[c - 1]);                            /* :475 */
            v += -tc * b[s - 1] * c4[c - 1];                                    /* :476 */
            v += -tc * b[s - 1] * (c5u * w1[c - 1]);                            /* :477 */
            Xm(X, s, c) = v;
        }
        for (int c = 1; c <= p; ++c) for (int s = 1; s <= lx - 1; ++s) {
            double v = Xm(X, s, c) + Xm(Xp, s + 1, c);                          /* :479 */
            { double z = 0.0; for (int d = 1; d <= r; ++d) z += Ym(Yp, s + 1, d) * (u1[d - 1] * w1[c - 1]); v += z; }     /* :480 */
            Xm(X, s, c) = v;
        }
        for (int c = 1; c <= p; ++c) for (int s = lb + 1; s <= lx; ++s) Xm(X, s, c) = 0.0;                               /* :482 */

        for (int c = 1; c <= r; ++c) for (int s = 1; s <= lb; ++s) {
            double v = -tc * b[s - 1] * kb[c - 1]; v += -tc * b[s - 1] * c2[c - 1]; v += -tc * b[s - 1] * c5[c - 1];     /* :485-487 */
            Ym(Y, s, c) = v;
        }
        for (int c = 1; c <= r; ++c) for (int s = 1; s <= lx - 1; ++s) Ym(Y, s, c) += Ym(Yp, s + 1, c);                  /* :489 */
        for (int c = 1; c <= r; ++c) for (int s = lb + 1; s <= lx; ++s) Ym(Y, s, c) = 0.0;                               /* :491 */

        for (int t = 1; t <= ldb - 1; ++t) for (int s = 1; s <= lb; ++s) Zm(Z, s, t) = -tc * b[s - 1] * dbar[t];         /* :494 */
        for (int t = 1; t <= ld1 - 1; ++t) for (int s = 1; s <= lb; ++s) {
            double v = Zm(Z, s, t);
            v += -tc * b[s - 1] * d1[t]; v += -tc * (b[s - 1] * c2u) * d1[t]; v += -tc * (b[s - 1] * c5u) * d1[t];       /* :496-498 */
            Zm(Z, s, t) = v;
        }
        for (int t = 1; t <= lm - 1; ++t) for (int s = 1; s <= lb; ++s) {
            double z = 0.0; for (int d = 1; d <= r; ++d) z += (b[s - 1] * c3[d - 1]) * Em(Ep, d, t + 1);
            Zm(Z, s, t) += -tc * z;                                             /* :500 */
        }
        for (int t = 1; t <= ld1 - 1; ++t) for (int s = 1; s <= lx - 1; ++s) {
            double yu = 0.0; for (int d = 1; d <= r; ++d) yu += Ym(Yp, s + 1, d) * u1[d - 1];
            Zm(Z, s, t) += yu * d1[t];                                          /* :502 */
        }
        for (int t = 1; t <= lm - 1; ++t) for (int s = 1; s <= lx - 1; ++s) Zm(Z, s, t) += Zm(Zp, s + 1, t + 1);         /* :504 */
        for (int t = 1; t <= lc6 - 1; ++t) for (int s = 1; s <= lb; ++s) Zm(Z, s, t) += -tc * b[s - 1] * c6[t];          /* :506 */
        for (int t = 1; t <= lm; ++t) for (int s = lb + 1; s <= lx; ++s) Zm(Z, s, t) = 0.0;                              /* :508 */
        for (int t = ldb; t <= lm; ++t) for (int s = 1; s <= lx; ++s) Zm(Z, s, t) = 0.0;                                 /* :510 */

        /* == update_upper_triangular_part! :514-562 (with the *_prev state) == */
        double beta[MAXR], alpha[MAXR], d[MAXB];
        for (int a = 1; a <= r; ++a) beta[a - 1] = -tc * kb[a - 1];                                                       /* :519 */
        for (int a = 1; a <= r; ++a) { double s = 0.0; for (int c = 1; c <= r; ++c) s += Km(Kp, c, a) * u1[c - 1]; beta[a - 1] += s; }   /* :520 */
        for (int a = 1; a <= r; ++a) { beta[a - 1] += -tc * c2[a - 1]; beta[a - 1] += -tc * c5[a - 1]; }                     /* :521-522 */
        if (lx > 0) for (int a = 1; a <= r; ++a) beta[a - 1] += Ym(Yp, 1, a);                                              /* :523-525 */
        for (int c = 1; c <= p; ++c) alpha[c - 1] = (1 - tc) * w1[c - 1];                                                  /* :528 */
        for (int c = 1; c <= p; ++c) alpha[c - 1] += -tc * f[c - 1];                                                       /* :529 */
        for (int c = 1; c <= p; ++c) { double s = 0.0; for (int a = 1; a <= r; ++a) s += Qm(Qp, a, c) * u1[a - 1]; alpha[c - 1] += s; }  /* :530 */
        for (int c = 1; c <= p; ++c) { alpha[c - 1] += -tc * c1[c - 1]; alpha[c - 1] += -tc * c4[c - 1]; }                   /* :531-532 */
        for (int c = 1; c <= p; ++c) { double s = 0.0; for (int a = 1; a <= r; ++a) s += (w1[c - 1] * u1[a - 1]) * kb[a - 1]; alpha[c - 1] += tc * s; }  /* :533 */
        if (lx > 0) for (int c = 1; c <= p; ++c) alpha[c - 1] += Xm(Xp, 1, c);                                             /* :534-536 */
        for (int c = 1; c <= p; ++c) { double s = 0.0; for (int a = 1; a <= r; ++a) s += M2(UW(j + 1), r, a, c) * beta[a - 1]; alpha[c - 1] += -1.0 * s; }  /* :538 */
        const int ld = ldb - 1;
        for (int t = 1; t <= ld; ++t) d[t - 1] = -tc * dbar[t];                                                           /* :540 */
        { double coef = 1 - tc + tc * ku; for (int t = 1; t <= ld1 - 1; ++t) d[t - 1] += coef * d1[t]; }                    /* :542 */
        { int hi = imin(ld, lm - 1);                                                                                      /* :544 */
          for (int t = 1; t <= hi; ++t) { double s = 0.0; for (int a = 1; a <= r; ++a) s += Em(Ep, a, t + 1) * (u1[a - 1] - tc * c3[a - 1]); d[t - 1] += s; } }
        for (int t = 1; t <= lc6 - 1; ++t) d[t - 1] += -tc * c6[t];                                                       /* :546 */
        if (lx > 0) { int hi = imin(ld, lm - 1); for (int t = 1; t <= hi; ++t) d[t - 1] += Zm(Zp, 1, t + 1); }              /* :547-550 */
        { int nc = imin(m, n - (j + 1));                                                                                  /* :551-553 */
          for (int t = 1; t <= nc; ++t) { double s = 0.0; for (int a = 1; a <= r; ++a) s += M2(DEX(j + 1), r, a, t) * beta[a - 1]; d[t - 1] += -1.0 * s; } }
        for (int c = 1; c <= p; ++c) Wg(&A, q1, c) = alpha[c - 1];                                                        /* :557 */
        for (int a = 1; a <= r; ++a) Wg(&A, q1, p + a) = beta[a - 1];                                                     /* :559 */
        for (int t = 1; t <= ld; ++t) *Bp(&A, q1, j + 1 + t) = d[t - 1];                                                  /* :561 */
        /* == update_lower_triangular_part! :564-571 == */
        *Bp(&A, q1, q1) = pivot_new;
        for (int s = 1; s <= lb; ++s) *Bp(&A, j + 1 + s, q1) = b[s - 1];
        for (int a = 1; a <= r; ++a) Vg(&A, q1, a) = kb[a - 1];
    }
    if (state) {
        double *o = state;
        memcpy(o, Q, sizeof(double) * r * p); o += r * p;
        memcpy(o, K, sizeof(double) * r * r); o += r * r;
        memcpy(o, E, sizeof(double) * r * lm); o += r * lm;
        memcpy(o, X, sizeof(double) * lx * p); o += lx * p;
        memcpy(o, Y, sizeof(double) * lx * r); o += lx * r;
        memcpy(o, Z, sizeof(double) * lx * lm);
    }
    /* ---- A.B[n,n] = A[n,n] :123, through getindex :50-102 with j = n-1 (k == i branch :84-91) ---- */
    if (nsteps == n - 1) {
        const int j = n - 1;
        bps_t Acur = { n, l, l + m, r, pr, Bf, (double *)U, Vf, Wf, Sf };
        double AtUn[MAXR];
        fast_UtA(&Acur, j + 1, AtUn);                         /* one column: (A'U)[1,:]  :54 */
        double UQ[MAXR], UK[MAXR], val = *Bp(&A, n, n), s1 = 0.0, s2 = 0.0;
        for (int c = 1; c <= p; ++c) { double s = 0.0; for (int a = 1; a <= r; ++a) s += Ug(&A, n, a) * Qm(Q, a, c); UQ[c - 1] = s; }
        for (int c = 1; c <= r; ++c) { double s = 0.0; for (int a = 1; a <= r; ++a) s += Ug(&A, n, a) * Km(K, a, c); UK[c - 1] = s; }
        for (int c = 1; c <= p; ++c) s1 += UQ[c - 1] * Sg(&A, n, c);
        for (int c = 1; c <= r; ++c) s2 += UK[c - 1] * AtUn[c - 1];
        val = val + s1 + s2;
        if (n <= j + l) {                                      /* :85-86 */
            double e = 0.0, x = 0.0, y = 0.0;
            for (int a = 1; a <= r; ++a) e += Ug(&A, n, a) * Em(E, a, 1);
            for (int c = 1; c <= p; ++c) x += Xm(X, 1, c) * Sg(&A, n, c);
            for (int c = 1; c <= r; ++c) y += Ym(Y, 1, c) * AtUn[c - 1];
            val = val + e + x + y + Zm(Z, 1, 1);
        } else if (n <= j + l + m) {                           /* :87-88 */
            double e = 0.0;
            for (int a = 1; a <= r; ++a) e += Ug(&A, n, a) * Em(E, a, 1);
            val = val + e;
        }
        *Bp(&A, n, n) = val;
    }
    free(AtU); free(UtU); free(uw); free(dex);
    return 0;
}

/* lmul!(Q', b) for QR{BPS}  (src/semiseparableqr.jl:863-919).  F = factors (B: (l, mw), V updated) */
int ssm_oracle_bps_lmul_qt(int n, int l, int mw, int r, const double *Bf, const double *U, const double *Vf, const double *tau, double *b) {
    bps_t F = { n, l, mw, r, 0, (double *)Bf, (double *)U, (double *)Vf, NULL, NULL };
    double *O = (double *)calloc((size_t)n * (r > 0 ? r : 1), sizeof(double));
    double *G = (double *)calloc((size_t)n * (l + 1), sizeof(double));
    double *Utb = (double *)calloc((size_t)(n + 2) * (r > 0 ? r : 1), sizeof(double));
    double *UtU = (double *)calloc((size_t)(n + 2) * r * r + 1, sizeof(double));
    double h[MAXR] = {0};
    {   /* tables :907-919, :573-585 */
        double cur[MAXR] = {0}, cu[MAXR * MAXR] = {0};
        for (int i = n; i >= 1; --i) {
            for (int a = 1; a <= r; ++a) cur[a - 1] += b[i - 1] * Ug(&F, i, a);
            memcpy(Utb + (size_t)i * r, cur, sizeof(double) * r);
            for (int c = 1; c <= r; ++c) for (int a = 1; a <= r; ++a) M2(cu, r, a, c) += Ug(&F, i, a) * Ug(&F, i, c);
            memcpy(UtU + (size_t)i * r * r, cu, sizeof(double) * r * r);
        }
    }
#define Gm(i, t) M2(G, n, i, t)
#define Om(i, a) M2(O, n, i, a)
    for (int j = 1; j <= n - 1; ++j) {
        int hi = imin(j + l, n);
        double c = 0.0, t0;
        t0 = 0.0; for (int a = 1; a <= r; ++a) t0 += Vg(&F, j, a) * Utb[(size_t)(j + 1) * r + a - 1]; c = t0;                 /* V[j,:]'Utb[j+1,:] */
        c += b[j - 1];
        t0 = 0.0; for (int i = j + 1; i <= hi; ++i) t0 += *Bp(&F, i, j) * b[i - 1]; c += t0;                                    /* B[jr,j]'b[jr] */
        { double tv[MAXR]; const double *Gj = UtU + (size_t)(j + 1) * r * r;                                                 /* V[j,:]'UtU[j+1]*h */
          for (int d = 1; d <= r; ++d) { double s = 0.0; for (int a = 1; a <= r; ++a) s += Vg(&F, j, a) * M2(Gj, r, a, d); tv[d - 1] = s; }
          t0 = 0.0; for (int d = 1; d <= r; ++d) t0 += tv[d - 1] * h[d - 1]; c += t0; }
        t0 = 0.0; for (int a = 1; a <= r; ++a) t0 += Ug(&F, j, a) * h[a - 1]; c += t0;                                         /* U[j,:]'h */
        { double tv[MAXR];                                                                                                   /* B[jr,j]'U[jr,:]*h */
          for (int a = 1; a <= r; ++a) { double s = 0.0; for (int i = j + 1; i <= hi; ++i) s += *Bp(&F, i, j) * Ug(&F, i, a); tv[a - 1] = s; }
          t0 = 0.0; for (int a = 1; a <= r; ++a) t0 += tv[a - 1] * h[a - 1]; c += t0; }
        { double tv[MAXB], sum = 0.0;                                                                                        /* sum(V[j,:]'U[jr,:]'G[jr,:]) */
          for (int i = j + 1; i <= hi; ++i) { double s = 0.0; for (int a = 1; a <= r; ++a) s += Vg(&F, j, a) * Ug(&F, i, a); tv[i - j - 1] = s; }
          for (int t = 1; t <= l + 1; ++t) { double s = 0.0; for (int i = j + 1; i <= hi; ++i) s += tv[i - j - 1] * Gm(i, t); sum += s; }
          c += sum; }
        { double sum = 0.0;                                                                                                  /* sum(G[j,:]' + B[jr,j]'G[jr,:]) */
          for (int t = 1; t <= l + 1; ++t) { double s = 0.0; for (int i = j + 1; i <= hi; ++i) s += *Bp(&F, i, j) * Gm(i, t); sum += Gm(j, t) + s; }
          c += sum; }
        for (int a = 1; a <= r; ++a) Om(j, a) = Ug(&F, j, a) * h[a - 1];                                                      /* :885 */
        for (int a = 1; a <= r; ++a) h[a - 1] += (-tau[j - 1] * c) * Vg(&F, j, a);                                            /* :887 */
        for (int t = 1; t <= imin(l + 1, n - j + 1); ++t)                                                                   /* :888-894 */
            Gm(j + t - 1, t) = (t == 1) ? -tau[j - 1] * c * 1.0 : -tau[j - 1] * c * *Bp(&F, j + t - 1, j);
    }
    for (int a = 1; a <= r; ++a) for (int i = 1; i <= n; ++i) b[i - 1] += Om(i, a);                                           /* :897-899 */
    for (int t = 1; t <= l + 1; ++t) for (int i = 1; i <= n; ++i) b[i - 1] += Gm(i, t);                                       /* :900-902 */
    { double s = 0.0; for (int a = 1; a <= r; ++a) s += Ug(&F, n, a) * h[a - 1]; b[n - 1] += s; }                             /* :904 */
    free(O); free(G); free(Utb); free(UtU);
    return 0;
}

int ssm_oracle_bps_upper_ldiv(int n, int l, int mw, int pw, const double *Bf, const double *Wf, const double *Sf, double *b) {
    bps_t F = { n, l, mw, 0, pw, (double *)Bf, NULL, NULL, (double *)Wf, (double *)Sf };
    bps_upper_ldiv(&F, b);
    return 0;
}
int ssm_oracle_bps_mul(int n, int l, int m, int r, int p, const double *B, const double *U, const double *V, const double *W,
                       const double *S, const double *x, double *y) {
    bps_t A = { n, l, m, r, p, (double *)B, (double *)U, (double *)V, (double *)W, (double *)S };
    bps_mul(&A, x, y);
    return 0;
}
int ssm_oracle_bps_dense(int n, int l, int m, int r, int p, const double *B, const double *U, const double *V, const double *W,
                         const double *S, double *out) {
    bps_t A = { n, l, m, r, p, (double *)B, (double *)U, (double *)V, (double *)W, (double *)S };
    for (int j = 1; j <= n; ++j) for (int k = 1; k <= n; ++k) out[(size_t)(k - 1) + (size_t)n * (size_t)(j - 1)] = bps_getindex0(&A, k, j);
    return 0;
}
double ssm_oracle_bps_relres(int n, int l, int m, int r, int p, const double *B, const double *U, const double *V, const double *W,
                             const double *S, const double *x, const double *b) {
    double *y = (double *)malloc(sizeof(double) * (size_t)n);
    ssm_oracle_bps_mul(n, l, m, r, p, B, U, V, W, S, x, y);
    double num = 0.0, den = 0.0;
    for (int i = 0; i < n; ++i) { double d = b[i] - y[i]; num += d * d; den += b[i] * b[i]; }
    free(y);
    return sqrt(num) / sqrt(den);
}

/* synthetic "dd" operands, reference layout */
int ssm_oracle_bps_generate(int n, int l, int m, int r, int p, int64_t s0, int64_t count, uint64_t seed, double *B, double *U, double *V,
                            double *W, double *S, double *b) {
    size_t bs = (size_t)(l + m + 1) * n, us = (size_t)n * r, ws = (size_t)n * p;
    const double su = 1.0 / sqrt((double)(r > 0 ? r : 1) * (double)n), sw = 1.0 / sqrt((double)(p > 0 ? p : 1) * (double)n);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < count; ++c) {
        uint64_t key = ssm_syskey(seed, (uint64_t)(s0 + c));
        if (B) for (int j = 0; j < n; ++j) for (int d = 0; d <= l + m; ++d) B[bs * c + (size_t)d + (size_t)(l + m + 1) * j] = ssm_bps_band_elem(key, n, l, m, d, j);
        for (int q = 0; q < r; ++q) for (int i = 0; i < n; ++i) {
            if (U) U[us * c + (size_t)i + (size_t)n * q] = ssm_bps_gen_elem(key, 1, n, r, i, q, su);
            if (V) V[us * c + (size_t)i + (size_t)n * q] = ssm_bps_gen_elem(key, 2, n, r, i, q, su);
        }
        for (int q = 0; q < p; ++q) for (int i = 0; i < n; ++i) {
            if (W) W[ws * c + (size_t)i + (size_t)n * q] = ssm_bps_gen_elem(key, 3, n, p, i, q, sw);
            if (S) S[ws * c + (size_t)i + (size_t)n * q] = ssm_bps_gen_elem(key, 4, n, p, i, q, sw);
        }
        if (b) for (int i = 0; i < n; ++i) b[(size_t)n * c + i] = ssm_bps_b_elem(key, i);
    }
    return 0;
}

/* batched drivers (OpenMP over systems): CPU baseline for config 3 */
int ssm_oracle_bps_qr_batch(int n, int l, int m, int r, int p, int64_t nb, const double *B, const double *U, const double *V,
                            const double *W, const double *S, double *Bf, double *Vf, double *Wf, double *Sf, double *tau) {
    size_t bs = (size_t)(l + m + 1) * n, us = (size_t)n * r, ws = (size_t)n * p, bfs = (size_t)(2 * l + m + 1) * n, wfs = (size_t)n * (p + r);
    int rc = 0;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < nb; ++s) {
        int e = ssm_oracle_bps_qr(n, l, m, r, p, B + bs * s, U + us * s, V + us * s, W + ws * s, S + ws * s, Bf + bfs * s, Vf + us * s,
                                  Wf + wfs * s, Sf + wfs * s, tau + (size_t)n * s);
        if (e) rc = e;
    }
    return rc;
}
int ssm_oracle_bps_solve_batch(int n, int l, int mw, int r, int pw, int64_t nb, const double *Bf, const double *U, const double *Vf,
                               const double *Wf, const double *Sf, const double *tau, double *b) {
    size_t us = (size_t)n * r, bfs = (size_t)(l + mw + 1) * n, wfs = (size_t)n * pw;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < nb; ++s) {
        ssm_oracle_bps_lmul_qt(n, l, mw, r, Bf + bfs * s, U + us * s, Vf + us * s, tau + (size_t)n * s, b + (size_t)n * s);
        ssm_oracle_bps_upper_ldiv(n, l, mw, pw, Bf + bfs * s, Wf + wfs * s, Sf + wfs * s, b + (size_t)n * s);
    }
    return 0;
}
