/* replay_julia_calls.c — replays, in plain C against include/ssm_b200.h only, the exact sequence of C-ABI calls every method override of
 * semiseparablematrices.jl_b200/julia/SemiseparableB200.jl makes (same functions, same argument order, stride 0, nb = 1), on small systems in
 * the reference memory layout, and checks each result against dense Gaussian elimination / dense Householder arithmetic done here.
 * There is no `julia` in the build image, so this is how the glue's call sequences get exercised on the GPU (tests/test_gpu_replay.py builds it
 * with gcc and runs it).  It also proves that the boundary is a C ABI: no CUDA header, no C++.
 *
 *   gcc -O1 -std=c11 -I include tests/c/replay_julia_calls.c -o replay -L <libdir> -lssmb200 -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ssm_b200.h"

#define CHECK(call)                                                                                         \
    do {                                                                                                    \
        int rc__ = (call);                                                                                  \
        if (rc__ != 0) { fprintf(stderr, "FAIL %s:%d: %s -> %d (%s)\n", __FILE__, __LINE__, #call, rc__, ssm_last_error()); exit(1); } \
    } while (0)
#define EXPECT(cond, ...)                                                                                   \
    do {                                                                                                    \
        if (!(cond)) { fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); exit(1); } \
    } while (0)

static unsigned long long rng_state = 88172645463325252ull;
static double rnd(void) {      /* xorshift, uniform in (-1, 1) */
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (double)(rng_state >> 11) / 9007199254740992.0 * 2.0 - 1.0;
}

/* dense n x n column-major helpers */
static void dense_solve(int n, const double *A, const double *b, double *x) {   /* Gaussian elimination with partial pivoting (row-major work copy) */
    double *M = malloc(sizeof(double) * (size_t)n * n);
    for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) M[(size_t)i * n + j] = A[i + (size_t)n * j];
    double *r = malloc(sizeof(double) * n);
    memcpy(r, b, sizeof(double) * n);
    for (int k = 0; k < n; ++k) {
        int p = k;
        for (int i = k + 1; i < n; ++i) if (fabs(M[(size_t)i * n + k]) > fabs(M[(size_t)p * n + k])) p = i;
        if (p != k) { for (int j = 0; j < n; ++j) { double t = M[(size_t)k * n + j]; M[(size_t)k * n + j] = M[(size_t)p * n + j]; M[(size_t)p * n + j] = t; } double t = r[k]; r[k] = r[p]; r[p] = t; }
        for (int i = k + 1; i < n; ++i) {
            double f = M[(size_t)i * n + k] / M[(size_t)k * n + k];
            if (f == 0.0) continue;
            for (int j = k; j < n; ++j) M[(size_t)i * n + j] -= f * M[(size_t)k * n + j];
            r[i] -= f * r[k];
        }
    }
    for (int i = n - 1; i >= 0; --i) { double s = r[i]; for (int j = i + 1; j < n; ++j) s -= M[(size_t)i * n + j] * x[j]; x[i] = s / M[(size_t)i * n + i]; }
    free(M); free(r);
}
static double relerr(int n, const double *x, const double *ref) {
    double a = 0, b = 0;
    for (int i = 0; i < n; ++i) { a += (x[i] - ref[i]) * (x[i] - ref[i]); b += ref[i] * ref[i]; }
    return sqrt(a) / sqrt(b);
}

/* AlmostBandedMatrix in the reference layout: bands (l+u+1) x n, A[k,j] at row u+k-j (0-based) of column j; U n x r; V r x n */
typedef struct { int n, l, u, r; double *bands, *U, *V; } AB;
static AB ab_random(int n, int l, int u, int r) {
    AB A = {n, l, u, r, calloc((size_t)(l + u + 1) * n, 8), calloc((size_t)n * (r ? r : 1), 8), calloc((size_t)n * (r ? r : 1), 8)};
    for (int j = 0; j < n; ++j)
        for (int d = 0; d <= l + u; ++d) { int k = j + d - u; if (k >= 0 && k < n) A.bands[d + (size_t)(l + u + 1) * j] = (k == j) ? (l + u + 2) + rnd() : rnd(); }
    for (int i = 0; i < n * r; ++i) { A.U[i] = rnd(); A.V[i] = rnd() / (r * n); }
    return A;
}
static void ab_dense(const AB *A, int n, double *D) {      /* leading n x n block (n <= A->n) */
    for (int j = 0; j < n; ++j)
        for (int k = 0; k < n; ++k) {
            double v;
            if (j <= k + A->u) { v = (k - j <= A->l) ? A->bands[(A->u + k - j) + (size_t)(A->l + A->u + 1) * j] : 0.0; }
            else { v = 0.0; for (int q = 0; q < A->r; ++q) v += A->U[k + (size_t)A->n * q] * A->V[q + (size_t)A->r * j]; }
            D[k + (size_t)n * j] = v;
        }
}

int main(void) {
    ssm_ctx *ctx = NULL;
    CHECK(ssm_ctx_create(0, &ctx));                                    /* context() */
    EXPECT(ssm_version() >= 100, "version");

    /* ---------------- qr(A) -> ldiv!(F, b)  (_almostbanded_qr, ldiv!(::QR, ::Vector)) ---------------- */
    const int n = 200, l = 2, u = 2, r = 2;
    AB A = ab_random(n, l, u, r);
    double *D = malloc(sizeof(double) * n * n), *b = malloc(8 * n), *x0 = malloc(8 * n), *x = malloc(8 * n);
    ab_dense(&A, n, D);
    for (int i = 0; i < n; ++i) b[i] = rnd();
    dense_solve(n, D, b, x0);
    ssm_ab *dA = NULL; ssm_abqr *dF = NULL; ssm_vec *v = NULL;
    CHECK(ssm_ab_create(ctx, n, l, u, r, 1, &dA));                      /* upload() */
    CHECK(ssm_ab_set(dA, A.bands, 0, A.U, 0, A.V, 0));
    CHECK(ssm_ab_qr_cols(dA, n - 1, &dF));                              /* device_qr!() */
    double *fdata = calloc((size_t)(2 * l + u + 1) * n, 8), *Uout = calloc((size_t)n * r, 8), *tau = calloc(n, 8);
    CHECK(ssm_abqr_get(dF, fdata, 0, Uout, 0, tau, 0));
    EXPECT(tau[n - 1] == 0.0, "tau[n] must be 0 (AlmostBandedMatrix.jl:137)");
    memcpy(x, b, 8 * n);
    CHECK(ssm_vec_create(ctx, n, 1, &v));                               /* device_vec() */
    CHECK(ssm_vec_set(v, x, 0));
    CHECK(ssm_abqr_ldiv(dF, v));                                        /* ldiv!(F, b) */
    CHECK(ssm_vec_get(v, x, 0));                                        /* fetch_vec!() */
    CHECK(ssm_vec_destroy(v));
    EXPECT(relerr(n, x, x0) <= 1e-12, "qr + ldiv!: relerr %.3e", relerr(n, x, x0));
    printf("qr + ldiv!                relerr %.2e\n", relerr(n, x, x0));

    /* ---------------- device_factor() of a QR this module did not make: ssm_abqr_set from F.factors / F.tau ---------------- */
    ssm_abqr *dF2 = NULL;
    CHECK(ssm_abqr_set(ctx, n, l, u, r, 1, fdata, 0, Uout, 0, A.V, 0, tau, 0, -1, &dF2));
    double *x2 = malloc(8 * n);
    memcpy(x2, b, 8 * n);
    CHECK(ssm_vec_create(ctx, n, 1, &v));
    CHECK(ssm_vec_set(v, x2, 0));
    CHECK(ssm_abqr_ldiv(dF2, v));
    CHECK(ssm_vec_get(v, x2, 0));
    CHECK(ssm_vec_destroy(v));
    EXPECT(memcmp(x, x2, 8 * n) == 0, "re-uploaded factors must solve bit for bit like the original");
    printf("ssm_abqr_set round trip   bit-identical\n");
    CHECK(ssm_abqr_destroy(dF2));

    /* ---------------- ldiv!(F, B::Matrix) ---------------- */
    {
        const int nrhs = 3;
        double *B = malloc(8 * (size_t)n * nrhs);
        for (int j = 0; j < nrhs; ++j) for (int i = 0; i < n; ++i) B[i + (size_t)n * j] = b[i] * (j + 1);
        ssm_mat *h = NULL;
        CHECK(ssm_mat_create(ctx, n, nrhs, 1, &h));
        CHECK(ssm_mat_set(h, B, 0));
        CHECK(ssm_abqr_ldiv_mat(dF, h));
        CHECK(ssm_mat_get(h, B, 0));
        CHECK(ssm_mat_destroy(h));
        for (int j = 0; j < nrhs; ++j) {
            double e = 0, s = 0;
            for (int i = 0; i < n; ++i) { double d = B[i + (size_t)n * j] - (j + 1) * x0[i]; e += d * d; s += x0[i] * x0[i] * (j + 1) * (j + 1); }
            EXPECT(sqrt(e / s) <= 1e-12, "ldiv!(F, B) column %d relerr %.3e", j, sqrt(e / s));
        }
        printf("ldiv!(F, B::Matrix)       ok\n");
        free(B);
    }

    /* ---------------- UpperTriangular(A) \ b and UnitUpperTriangular(A) \ b  (materialize!) ---------------- */
    for (int unit = 0; unit <= 1; ++unit) {
        double *T = malloc(sizeof(double) * n * n);
        memcpy(T, D, sizeof(double) * n * n);
        for (int j = 0; j < n; ++j) for (int k = j + 1; k < n; ++k) T[k + (size_t)n * j] = 0.0;
        if (unit) for (int j = 0; j < n; ++j) T[j + (size_t)n * j] = 1.0;
        dense_solve(n, T, b, x2);
        memcpy(x, b, 8 * n);
        CHECK(ssm_vec_create(ctx, n, 1, &v));
        CHECK(ssm_vec_set(v, x, 0));
        CHECK(ssm_ab_tri_ldiv(dA, v, 1, unit));
        CHECK(ssm_vec_get(v, x, 0));
        CHECK(ssm_vec_destroy(v));
        EXPECT(relerr(n, x, x2) <= 1e-12, "triangular (unit=%d) relerr %.3e", unit, relerr(n, x, x2));
        free(T);
    }
    printf("(Unit)UpperTriangular \\ b ok\n");
    CHECK(ssm_abqr_destroy(dF));
    CHECK(ssm_ab_destroy(dA));

    /* ---------------- A \ b for one larger system: chunked path (Base.:\) ---------------- */
    {
        const int N = 1500;
        AB L = ab_random(N, 1, 0, 1);
        double *DL = malloc(sizeof(double) * N * N), *bl = malloc(8 * N), *xl = malloc(8 * N), *xr = malloc(8 * N);
        ab_dense(&L, N, DL);
        for (int i = 0; i < N; ++i) bl[i] = rnd();
        dense_solve(N, DL, bl, xr);
        ssm_abc *h = NULL;
        CHECK(ssm_abc_create(ctx, N, 1, 0, 1, &h));
        CHECK(ssm_abc_set(h, L.bands, L.U, L.V));
        CHECK(ssm_abc_solve(h, bl, xl, 0));
        CHECK(ssm_abc_destroy(h));
        EXPECT(relerr(N, xl, xr) <= 1e-12, "chunked A \\ b relerr %.3e", relerr(N, xl, xr));
        printf("A \\ b (chunked, n=%d)   relerr %.2e\n", N, relerr(N, xl, xr));
        free(DL); free(bl); free(xl); free(xr); free(L.bands); free(L.U); free(L.V);
    }

    /* ---------------- partial_qr / grow_qr! / continue_qr!  (adaptive QR across a growth step) ---------------- */
    {
        const int n0 = 120, k0 = n0 - l, k1 = n - 1;       /* factor as close to the edge as growth allows */
        /* operands of the leading n0 x n0 block, in arrays of their own size */
        double *b0 = calloc((size_t)(l + u + 1) * n0, 8), *U0 = calloc((size_t)n0 * r, 8), *V0 = calloc((size_t)n0 * r, 8);
        for (int j = 0; j < n0; ++j) for (int d = 0; d <= l + u; ++d) { int k = j + d - u; if (k >= 0 && k < n0) b0[d + (size_t)(l + u + 1) * j] = A.bands[d + (size_t)(l + u + 1) * j]; }
        for (int q = 0; q < r; ++q) for (int i = 0; i < n0; ++i) U0[i + (size_t)n0 * q] = A.U[i + (size_t)n * q];
        for (int j = 0; j < n0; ++j) for (int q = 0; q < r; ++q) V0[q + (size_t)r * j] = A.V[q + (size_t)r * j];
        ssm_ab *d0 = NULL, *d1 = NULL; ssm_abqr *G = NULL, *H = NULL;
        CHECK(ssm_ab_create(ctx, n0, l, u, r, 1, &d0));
        CHECK(ssm_ab_set(d0, b0, 0, U0, 0, V0, 0));
        CHECK(ssm_ab_qr_cols(d0, k0, &G));                              /* partial_qr(A0, k0) */
        CHECK(ssm_ab_resize(d0, n, &d1));                               /* grow_qr!: resize + set_rows + abqr_resize */
        CHECK(ssm_ab_set_rows(d1, n0 - u > 0 ? n0 - u : 0, n, A.bands, 0, A.U, 0, A.V, 0));
        CHECK(ssm_abqr_resize(G, d1));
        CHECK(ssm_abqr_continue(G, d1, k1));                            /* continue_qr!(F, n - 1) */
        double *f1 = calloc((size_t)(2 * l + u + 1) * n, 8), *u1 = calloc((size_t)n * r, 8), *t1 = calloc(n, 8);
        CHECK(ssm_abqr_get(G, f1, 0, u1, 0, t1, 0));
        double e = 0;
        for (int i = 0; i < (2 * l + u + 1) * n; ++i) e = fmax(e, fabs(f1[i] - fdata[i]));
        for (int i = 0; i < n; ++i) e = fmax(e, fabs(t1[i] - tau[i]));
        EXPECT(e <= 1e-12, "grown + continued factors differ from the one-shot qr by %.3e", e);
        printf("partial/grow/continue     max |dF| %.2e\n", e);
        /* refusing to grow reflectors that saw a truncated column */
        CHECK(ssm_ab_qr_cols(d0, n0 - 1, &H));
        EXPECT(ssm_abqr_resize(H, d1) == SSM_E_STATE, "growth of a complete factorisation must be SSM_E_STATE");
        CHECK(ssm_abqr_destroy(G)); CHECK(ssm_abqr_destroy(H)); CHECK(ssm_ab_destroy(d0)); CHECK(ssm_ab_destroy(d1));
        free(b0); free(U0); free(V0); free(f1); free(u1); free(t1);
    }

    /* ---------------- BandedPlusSemiseparableMatrix: qr(A), lmul!(F.Q', b), ldiv!(F.R, b) ---------------- */
    {
        const int nb_ = 90, bl_ = 2, bm = 2, br = 2, bp = 2;
        const int w = bl_ + bm + 1;
        double *B = calloc((size_t)w * nb_, 8), *U = malloc(8 * nb_ * br), *V = malloc(8 * nb_ * br), *W = malloc(8 * nb_ * bp), *S = malloc(8 * nb_ * bp);
        for (int j = 0; j < nb_; ++j) for (int d = 0; d < w; ++d) { int k = j + d - bm; if (k >= 0 && k < nb_) B[d + (size_t)w * j] = (k == j) ? w + 1 + rnd() : rnd(); }
        for (int i = 0; i < nb_ * br; ++i) { U[i] = rnd() / sqrt(br * nb_); V[i] = rnd() / sqrt(br * nb_); }
        for (int i = 0; i < nb_ * bp; ++i) { W[i] = rnd() / sqrt(bp * nb_); S[i] = rnd() / sqrt(bp * nb_); }
        double *Dd = malloc(sizeof(double) * nb_ * nb_), *bb = malloc(8 * nb_), *xr = malloc(8 * nb_), *xb = malloc(8 * nb_);
        for (int j = 0; j < nb_; ++j) for (int k = 0; k < nb_; ++k) {      /* B + tril(UV',-1) + triu(WS',1) */
            double val = (abs(k - j) <= (k > j ? bl_ : bm)) ? B[(bm + k - j) + (size_t)w * j] : 0.0;
            if (k > j) for (int q = 0; q < br; ++q) val += U[k + (size_t)nb_ * q] * V[j + (size_t)nb_ * q];
            if (k < j) for (int q = 0; q < bp; ++q) val += W[k + (size_t)nb_ * q] * S[j + (size_t)nb_ * q];
            Dd[k + (size_t)nb_ * j] = val;
        }
        for (int i = 0; i < nb_; ++i) bb[i] = rnd();
        dense_solve(nb_, Dd, bb, xr);
        ssm_bps *dB = NULL; ssm_bpsqr *dQ = NULL;
        CHECK(ssm_bps_create(ctx, nb_, bl_, bm, br, bp, 1, &dB));       /* upload(A) */
        CHECK(ssm_bps_set(dB, B, 0, U, V, 0, W, S, 0));
        CHECK(ssm_bps_qr(dB, &dQ));                                     /* qr(A) */
        double *Bf = calloc((size_t)(2 * bl_ + bm + 1) * nb_, 8), *Uf = malloc(8 * nb_ * br), *Vf = malloc(8 * nb_ * br),
               *Wf = malloc(8 * nb_ * (bp + br)), *Sf = malloc(8 * nb_ * (bp + br)), *tf = malloc(8 * nb_);
        CHECK(ssm_bpsqr_get(dQ, Bf, 0, Uf, Vf, 0, Wf, Sf, 0, tf, 0));
        memcpy(xb, bb, 8 * nb_);
        CHECK(ssm_vec_create(ctx, nb_, 1, &v));
        CHECK(ssm_vec_set(v, xb, 0));
        CHECK(ssm_bpsqr_lmul_qt(dQ, v));                                /* lmul!(F.Q', b) */
        CHECK(ssm_vec_get(v, xb, 0));
        CHECK(ssm_vec_destroy(v));
        double nb2 = 0, nq2 = 0;
        for (int i = 0; i < nb_; ++i) { nb2 += bb[i] * bb[i]; nq2 += xb[i] * xb[i]; }
        EXPECT(fabs(sqrt(nq2) - sqrt(nb2)) <= 1e-13 * sqrt(nb2), "Q' must preserve the 2-norm");
        CHECK(ssm_vec_create(ctx, nb_, 1, &v));
        CHECK(ssm_vec_set(v, xb, 0));
        CHECK(ssm_bpsqr_upper_ldiv(dQ, v));                             /* ldiv!(F.R, b) */
        CHECK(ssm_vec_get(v, xb, 0));
        CHECK(ssm_vec_destroy(v));
        EXPECT(relerr(nb_, xb, xr) <= 1e-12, "BPS qr + Q'b + R\\b relerr %.3e", relerr(nb_, xb, xr));
        printf("BPS qr, Q'b, R \\ b        relerr %.2e\n", relerr(nb_, xb, xr));
        /* ldiv!(UpperTriangular(A), b) on an UNFACTORED matrix (test_bandedplussemiseparable.jl:17): the operand itself is uploaded */
        for (int j = 0; j < nb_; ++j) for (int k = j + 1; k < nb_; ++k) Dd[k + (size_t)nb_ * j] = 0.0;
        dense_solve(nb_, Dd, bb, xr);
        memcpy(xb, bb, 8 * nb_);
        CHECK(ssm_vec_create(ctx, nb_, 1, &v));
        CHECK(ssm_vec_set(v, xb, 0));
        CHECK(ssm_bps_upper_ldiv(dB, v));
        CHECK(ssm_vec_get(v, xb, 0));
        CHECK(ssm_vec_destroy(v));
        EXPECT(relerr(nb_, xb, xr) <= 1e-12, "UpperTriangular(BPS) \\ b relerr %.3e", relerr(nb_, xb, xr));
        CHECK(ssm_bpsqr_destroy(dQ)); CHECK(ssm_bps_destroy(dB));
        free(B); free(U); free(V); free(W); free(S); free(Dd); free(bb); free(xr); free(xb); free(Bf); free(Uf); free(Vf); free(Wf); free(Sf); free(tf);
    }

    /* ---------------- solve_batch: ssm_ab_solve_host on a few systems ---------------- */
    {
        const int nbat = 40, nn = 64;
        size_t eb = (size_t)(l + u + 1) * nn, eu = (size_t)nn * r;
        double *bands = malloc(8 * eb * nbat), *U = malloc(8 * eu * nbat), *V = malloc(8 * eu * nbat), *rb = malloc(8 * (size_t)nn * nbat), *xs = malloc(8 * (size_t)nn * nbat);
        double *Dd = malloc(sizeof(double) * nn * nn), *xr = malloc(8 * nn);
        AB *sys = malloc(sizeof(AB) * nbat);
        for (int s = 0; s < nbat; ++s) {
            sys[s] = ab_random(nn, l, u, r);
            memcpy(bands + eb * s, sys[s].bands, 8 * eb); memcpy(U + eu * s, sys[s].U, 8 * eu); memcpy(V + eu * s, sys[s].V, 8 * eu);
            for (int i = 0; i < nn; ++i) rb[(size_t)nn * s + i] = rnd();
        }
        CHECK(ssm_ab_solve_host(ctx, nn, l, u, r, nbat, bands, U, V, rb, xs));
        for (int s = 0; s < nbat; ++s) {
            ab_dense(&sys[s], nn, Dd);
            dense_solve(nn, Dd, rb + (size_t)nn * s, xr);
            EXPECT(relerr(nn, xs + (size_t)nn * s, xr) <= 1e-12, "solve_batch system %d relerr %.3e", s, relerr(nn, xs + (size_t)nn * s, xr));
            free(sys[s].bands); free(sys[s].U); free(sys[s].V);
        }
        printf("solve_batch (host path)   ok\n");
        free(bands); free(U); free(V); free(rb); free(xs); free(Dd); free(xr); free(sys);
    }

    CHECK(ssm_ctx_destroy(ctx));
    free(D); free(b); free(x0); free(x); free(x2); free(fdata); free(Uout); free(tau); free(A.bands); free(A.U); free(A.V);
    printf("replay OK\n");
    return 0;
}
