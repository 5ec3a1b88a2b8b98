// bps_emul.cpp — DEBUGGING HARNESS (see ab_emul.cpp): compiles csrc/bps_core.cuh for the host and runs one system
// through exactly the record layout / step order the CUDA kernels use, so the streaming restatement of the BPS
// sweep can be checked against the oracle without a GPU.  Never shipped, never loaded by the product.
#include <cstdint>
#include <vector>
#include "../../semiseparablematrices.jl_b200/csrc/bps_core.cuh"

using namespace ssm;

template <int L, int M, int R, int P>
static int run(int n, int l, int m, int r, int p, const double *B, const double *U, const double *V, const double *W, const double *S,
               const double *b, double *Bf, double *Vf, double *Wf, double *Sf, double *tau, double *c, double *x) {
    using D = BpsDims<L, M, R, P>;
    constexpr int FRONT = 8, NFA = D::NFA, NFO = D::NFO, LM = L + M;
    const int np = FRONT + n + 48;
    std::vector<double> rec((size_t)np * NFA, 0.0), out((size_t)np * NFO, 0.0);
    // pack (zero padded to the instantiated shape): B0[i, i-L+d] = B[(m + i - j) + (l+m+1) j]
    for (int i = 0; i < n; ++i) {
        double *rr = &rec[(size_t)(FRONT + i) * NFA];
        for (int d = 0; d <= LM; ++d) { int j = i - L + d; if (j >= 0 && j < n && j - i <= m && i - j <= l) rr[d] = B[(m + i - j) + (size_t)(l + m + 1) * j]; }
        for (int a = 0; a < r; ++a) { rr[D::FU + a] = U[i + (size_t)n * a]; rr[D::FV + a] = V[i + (size_t)n * a]; }
        for (int q = 0; q < p; ++q) { rr[D::FW + q] = W[i + (size_t)n * q]; rr[D::FS + q] = S[i + (size_t)n * q]; }
    }
    double gtot[R * R] = {0};
    for (int i = 0; i < n; ++i) for (int a = 0; a < R; ++a) for (int cc = 0; cc < R; ++cc) gtot[a * R + cc] = fma(rec[(size_t)(FRONT + i) * NFA + D::FU + a], rec[(size_t)(FRONT + i) * NFA + D::FU + cc], gtot[a * R + cc]);
    BpsState<L, M, R, P> st;
    bps_init<L, M, R, P>(st, gtot);
    for (int q = 0; q < n; ++q)
        bps_step<L, M, R, P>(st, [&](int rel, int f) { return rec[(size_t)(FRONT + q + rel) * NFA + f]; }, q == n - 1, &out[(size_t)q * NFO]);
    // unpack into the reference layout of the ORIGINAL shape (l, m, r, p): Bf (l, l+m), Vf n x r, Wf/Sf n x (p+r)
    const int mw = l + m, nd = 2 * l + m + 1;
    for (int q = 0; q < n; ++q) {
        const double *o = &out[(size_t)q * NFO];
        for (int t = 0; t <= mw; ++t) if (q + t < n) Bf[(mw - t) + (size_t)nd * (q + t)] = o[D::OD + t];
        for (int s2 = 1; s2 <= l; ++s2) if (q + s2 < n) Bf[(mw + s2) + (size_t)nd * q] = o[D::OB + s2 - 1];
        for (int a = 0; a < r; ++a) Vf[q + (size_t)n * a] = o[D::OV + a];
        for (int q2 = 0; q2 < p; ++q2) { Wf[q + (size_t)n * q2] = o[D::OWA + q2]; Sf[q + (size_t)n * q2] = S[q + (size_t)n * q2]; }
        for (int a = 0; a < r; ++a) { Wf[q + (size_t)n * (p + a)] = o[D::OWB + a]; Sf[q + (size_t)n * (p + a)] = o[D::OSU + a]; }
        tau[q] = o[D::OT];
        // padded parts must vanish
        for (int t = mw + 1; t <= LM; ++t) if (o[D::OD + t] != 0.0 && q + t < n) return 200;
    }
    // Q'b
    double ttot[R] = {0};
    for (int i = 0; i < n; ++i) for (int a = 0; a < R; ++a) ttot[a] = fma(rec[(size_t)(FRONT + i) * NFA + D::FU + a], b[i], ttot[a]);
    std::vector<double> rhs(np, 0.0);
    for (int i = 0; i < n; ++i) rhs[i] = b[i];
    BpsQtState<L, R> qs;
    bps_qt_init<L, R>(qs, gtot, ttot, [&](int row, int a) { return rec[(size_t)(FRONT + row) * NFA + D::FU + a]; }, [&](int row) { return rhs[row]; });
    for (int q = 0; q < n; ++q) {
        const double *o = &out[(size_t)q * NFO];
        double cc = bps_qt_step<L, R>(qs, [&](int rel, int a) { return rec[(size_t)(FRONT + q + rel) * NFA + D::FU + a]; }, o + D::OV, o + D::OB, o[D::OT], rhs[q + L], q == n - 1);
        c[q] = cc;
    }
    // back substitution with W = [alpha beta], S = [S A'U]
    BpsBsState<LM, P + R> bs;
    bps_bs_init<LM, P + R>(bs);
    for (int j = n - 1; j >= 0; --j) {
        const double *o = &out[(size_t)j * NFO];
        double w[P + R], s2[P + R];
        for (int q2 = 0; q2 < P; ++q2) { w[q2] = o[D::OWA + q2]; s2[q2] = rec[(size_t)(FRONT + j) * NFA + D::FS + q2]; }
        for (int a = 0; a < R; ++a) { w[P + a] = o[D::OWB + a]; s2[P + a] = o[D::OSU + a]; }
        x[j] = bps_bs_step<LM, P + R>(bs, o + D::OD, w, s2, c[j]);
    }
    return 0;
}

extern "C" int emul_bps(int n, int l, int m, int r, int p, const double *B, const double *U, const double *V, const double *W,
                        const double *S, const double *b, double *Bf, double *Vf, double *Wf, double *Sf, double *tau, double *c, double *x) {
#define X(L_, M_, R_, P_) if (l <= L_ && m <= M_ && r <= R_ && p <= P_) return run<L_, M_, R_, P_>(n, l, m, r, p, B, U, V, W, S, b, Bf, Vf, Wf, Sf, tau, c, x);
    X(1, 1, 1, 1) X(2, 2, 2, 2) X(3, 3, 2, 2) X(4, 5, 3, 3)
#undef X
    return -3;
}
