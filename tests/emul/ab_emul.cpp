// ab_emul.cpp — DEBUGGING HARNESS, never shipped and never loaded by the product.
// Compiles the per-system device arithmetic (csrc/ab_core.cuh) for the HOST so that the window / slot
// bookkeeping can be checked against the oracle in the CPU-only test tier (no GPU in the dev container).
// The CUDA kernels run the very same templates; record packing mirrors csrc/util_kernels.cu.
#include <cstdint>
#include <vector>
#include "../../semiseparablematrices.jl_b200/csrc/ab_core.cuh"

using namespace ssm;

template <int L, int U, int R>
static int run(int n, const double *bands, const double *Uin, const double *Vin, const double *b, double *factors,
               double *Uout, double *tau, double *c, double *x) {
    using D = AbDims<L, U, R>;
    constexpr int P = D::P, UB = D::UB, NFA = D::NFA, NFR = D::NFR, NFQ = D::NFQ;
    const int np = n + 48;
    std::vector<double> AU((size_t)np * NFA, 0.0), V((size_t)np * (R > 0 ? R : 1), 0.0), rhs(np, 0.0), RU((size_t)np * NFR, 0.0), QT((size_t)np * NFQ, 0.0);
    for (int i = 0; i < n; ++i) {
        for (int d = 0; d <= L + U; ++d) { int j = i - L + d; if (j >= 0 && j < n) AU[(size_t)i * NFA + d] = bands[(L + U - d) + (L + U + 1) * (size_t)j]; }
        for (int q = 0; q < R; ++q) AU[(size_t)i * NFA + L + U + 1 + q] = Uin[i + (size_t)n * q];
        for (int q = 0; q < R; ++q) V[(size_t)i * R + q] = Vin[q + (size_t)R * i];
        rhs[i] = b ? b[i] : 0.0;
    }
    AbState<L, U, R> st;
    ab_prologue<L, U, R, true>(st, [&](int rho, int f) { return AU[(size_t)rho * NFA + f]; },
                               [&](int cc, int q) { return V[(size_t)cc * R + q]; }, [&](int rho) { return rhs[rho]; });
    for (int kb = 0; kb < n; kb += P)
        for (int ph = 0; ph < P; ++ph) {
            int k = kb + ph;
            double out_r[NFR], out_q[NFQ], out_c;
            ab_step<L, U, R, true>(st, ph, &AU[(size_t)(k + L) * NFA], &V[(size_t)(k + UB + 1) * (R > 0 ? R : 1)], rhs[k + L], k < n - 1, out_r, out_q, out_c);
            if (k < n) {
                for (int f = 0; f < NFR; ++f) RU[(size_t)k * NFR + f] = out_r[f];
                for (int f = 0; f < NFQ; ++f) QT[(size_t)k * NFQ + f] = out_q[f];
                rhs[k] = out_c;
            }
        }
    // unpack (mirrors abqr_unpack_kernel)
    const int nd = 2 * L + U + 1;
    for (int k = 0; k < n; ++k) {
        for (int t = 0; t <= UB; ++t) if (k + t < n) factors[(UB - t) + (size_t)nd * (k + t)] = RU[(size_t)k * NFR + t];
        for (int i = 1; i <= L; ++i) if (k + i < n) factors[(UB + i) + (size_t)nd * k] = QT[(size_t)k * NFQ + i - 1];
        for (int q = 0; q < R; ++q) Uout[k + (size_t)n * q] = RU[(size_t)k * NFR + UB + 1 + q];
        tau[k] = QT[(size_t)k * NFQ + L];
        c[k] = rhs[k];
    }
    // separate Q'b pass from stored reflectors must reproduce c
    {
        std::vector<double> r2(np, 0.0);
        for (int i = 0; i < n; ++i) r2[i] = b ? b[i] : 0.0;
        QtState<L> qs;
        qt_prologue<L>(qs, [&](int rho) { return r2[rho]; });
        for (int kb = 0; kb < n; kb += P)
            for (int ph = 0; ph < P; ++ph) {
                int k = kb + ph;
                double cc = qt_step<L>(qs, ph, &QT[(size_t)k * NFQ], r2[k + L]);
                if (k < n) { r2[k] = cc; if (cc != rhs[k]) return 100 + k; }
            }
    }
    // back substitution
    BsState<UB, R> bs;
    bs_init<UB, R>(bs);
    constexpr int PX = UB + 1;
    const int nblk = (n + PX - 1) / PX;
    for (int ib = (nblk - 1) * PX; ib >= 0; ib -= PX)
        for (int ph = PX - 1; ph >= 0; --ph) {
            int i = ib + ph;
            if (i < n) { double xv = bs_step<UB, R>(bs, ph, &RU[(size_t)i * NFR], &V[(size_t)(i + UB + 1) * (R > 0 ? R : 1)], rhs[i]); rhs[i] = xv; }
        }
    for (int i = 0; i < n; ++i) x[i] = rhs[i];
    return 0;
}

extern "C" int emul_ab(int n, int l, int u, int r, const double *bands, const double *U, const double *V, const double *b,
                       double *factors, double *Uout, double *tau, double *c, double *x) {
#define X(L_, U_, R_) if (l == L_ && u == U_ && r == R_) return run<L_, U_, R_>(n, bands, U, V, b, factors, Uout, tau, c, x);
    X(1, 0, 1) X(1, 1, 1) X(1, 0, 2) X(2, 0, 2) X(2, 1, 2) X(2, 2, 1) X(2, 2, 2) X(3, 3, 2) X(3, 3, 3) X(4, 4, 2) X(0, 2, 1) X(3, 1, 3)
#undef X
    return -3;
}
