// ab_generic.cuh — runtime-(l,u,r) device routines for any AlmostBanded shape within SSM_MAX_BW / SSM_MAX_RANK.
// Same arithmetic as ab_core.cuh with an explicitly shifted window in local memory: the completeness path for
// shapes without a register-resident instance, and the home of the rarely used entry points
// (Q*b, UnitUpperTriangular / LowerTriangular solves — AlmostBandedMatrix.jl:201-207,250-270).
#pragma once
#include "ssm_internal.cuh"

namespace ssm {

// kstart > 0 continues a partial factorisation in place (ssm_abqr_continue): reflectors 0 .. kstart-1 exist, the partially updated rows
// kstart .. kstart+l-1 come back out of the factor records (RU: diagonal and right of it, U row; QT of columns kstart ..: the entries below).
__device__ inline void generic_ab_qr(const AbArgs &a, int l, int u, int r, bool with_rhs, bool store_q, int64_t tile, int lane, int kstart = 0) {
    constexpr int MB = SSM_MAX_BW, MR = SSM_MAX_RANK;
    const int ub = l + u, nfa = l + u + 1 + r, nfr = nfa, nfq = l + 1, n = a.n;
    const double *au = a.AU + tile * a.np * nfa * 32 + lane;
    const double *vv = a.V + tile * a.np * (r > 0 ? r : 1) * 32 + lane;
    double *rhs = with_rhs ? a.rhs + tile * a.np * 32 + lane : nullptr;
    const double *rin = with_rhs ? a.rhs_in + tile * a.np * 32 + lane : nullptr;
    double *ru = a.RU + tile * a.np * nfr * 32 + lane;
    double *qt = store_q ? a.QT + tile * a.np * nfq * 32 + lane : nullptr;
    double W[MB + 1][2 * MB + 1], UC[MB + 1][MR], RH[MB + 1], v[MB];
    for (int i = 0; i <= l; ++i) { for (int t = 0; t <= ub; ++t) W[i][t] = 0.0; for (int q = 0; q < MR; ++q) UC[i][q] = 0.0; RH[i] = 0.0; }
    if (kstart == 0) {
    for (int rho = 0; rho < l; ++rho) {       // rows resident before step 0 (AlmostBandedMatrix.jl:30-38 for their fill part)
        for (int q = 0; q < r; ++q) UC[rho][q] = au[((int64_t)rho * nfa + l + u + 1 + q) * 32];
        RH[rho] = with_rhs ? rin[(int64_t)rho * 32] : 0.0;
        for (int t = 0; t <= ub; ++t) {
            if (t - rho <= u) W[rho][t] = au[((int64_t)rho * nfa + (t - rho + l)) * 32];
            else { double s = 0.0; for (int q = 0; q < r; ++q) s = fma(UC[rho][q], vv[((int64_t)t * r + q) * 32], s); W[rho][t] = s; }
        }
    }
    } else {
    for (int i = 0; i < l; ++i) {             // window row i = matrix row kstart + i, window column t = matrix column kstart + t
        const int64_t rho = (int64_t)kstart + i;
        for (int q = 0; q < r; ++q) UC[i][q] = ru[(rho * nfr + ub + 1 + q) * 32];
        for (int t = 0; t <= ub; ++t)
            W[i][t] = (t >= i) ? ru[(rho * nfr + (t - i)) * 32] : qt[(((int64_t)kstart + t) * nfq + (i - t - 1)) * 32];
    }
    }
    for (int k = kstart; k < n; ++k) {
        const int64_t e = k + l;                 // entering row
        for (int t = 0; t <= ub; ++t) W[l][t] = (t <= l + u) ? au[(e * nfa + t) * 32] : 0.0;
        for (int q = 0; q < r; ++q) UC[l][q] = au[(e * nfa + l + u + 1 + q) * 32];
        RH[l] = with_rhs ? rin[e * 32] : 0.0;
        const double x0 = W[0][0];
        double ssq = x0 * x0;
        for (int i = 1; i <= l; ++i) ssq = fma(W[i][0], W[i][0], ssq);
        const double nu = copysign(sqrt(ssq), x0), xi1 = x0 + nu;
        const bool nz = (k < a.ncols) && (ssq != 0.0);
        const double inv_xi = nz ? 1.0 / xi1 : 0.0, tau = nz ? xi1 / nu : 0.0;
        for (int i = 1; i <= l; ++i) v[i - 1] = W[i][0] * inv_xi;
        double tail[SSM_MAX_BW];      // what the factor storage holds below the diagonal: v, or the untouched column
        for (int i = 1; i <= l; ++i) tail[i - 1] = nz ? v[i - 1] : W[i][0];
        W[0][0] = nz ? -nu : x0;
        for (int t = 1; t <= ub; ++t) {
            double s = W[0][t];
            for (int i = 1; i <= l; ++i) s = fma(v[i - 1], W[i][t], s);
            s *= tau; W[0][t] -= s;
            for (int i = 1; i <= l; ++i) W[i][t] = fma(-v[i - 1], s, W[i][t]);
        }
        for (int q = 0; q < r; ++q) {
            double s = UC[0][q];
            for (int i = 1; i <= l; ++i) s = fma(v[i - 1], UC[i][q], s);
            s *= tau; UC[0][q] -= s;
            for (int i = 1; i <= l; ++i) UC[i][q] = fma(-v[i - 1], s, UC[i][q]);
        }
        if (with_rhs) {
            double s = RH[0];
            for (int i = 1; i <= l; ++i) s = fma(v[i - 1], RH[i], s);
            s *= tau; RH[0] -= s;
            for (int i = 1; i <= l; ++i) RH[i] = fma(-v[i - 1], s, RH[i]);
        }
        for (int t = 0; t <= ub; ++t) ru[((int64_t)k * nfr + t) * 32] = W[0][t];
        for (int q = 0; q < r; ++q) ru[((int64_t)k * nfr + ub + 1 + q) * 32] = UC[0][q];
        if (store_q) { for (int i = 1; i <= l; ++i) qt[((int64_t)k * nfq + i - 1) * 32] = tail[i - 1]; qt[((int64_t)k * nfq + l) * 32] = tau; }
        if (with_rhs) rhs[(int64_t)k * 32] = RH[0];
        const int64_t c = (int64_t)k + 1 + ub;   // column gained by the surviving rows (fill region)
        for (int i = 1; i <= l; ++i) {
            for (int t = 1; t <= ub; ++t) W[i - 1][t - 1] = W[i][t];
            double s = 0.0;
            for (int q = 0; q < r; ++q) s = fma(UC[i][q], vv[(c * r + q) * 32], s);
            W[i - 1][ub] = s;
            for (int q = 0; q < r; ++q) UC[i - 1][q] = UC[i][q];
            RH[i - 1] = RH[i];
        }
        if (l == 0) { /* nothing survives */ }
    }
}

// Growth fix-up of a partial factorisation (ssm_abqr_resize; what adaptive QR needs when the cached operand of AlmostBandedMatrix.jl:357-394
// has grown under a factorisation in progress).  a.RU / a.QT hold k0 reflectors formed when the operand had n0 rows and columns; a.AU / a.V are
// the GROWN operand (n = a.n >= n0, old part unchanged).  Explicit entries of the factor records in the NEW columns c >= n0 — they can occur in
// rows j0 .. k0+l-1, j0 = max(0, n0 - 2l - u) — were formed when those columns did not exist.  They are re-derived by replaying steps
// j0 .. k0-1 of the sweep with the STORED reflectors (lmul!(Q', .) restricted to the new columns, :201-207):
//   (1) backwards over k = k0-1 .. j0, un-applying H_k (an involution) to the window of U rows: recovers the U rows as they were before step j0;
//   (2) forwards, the sweep's own bookkeeping (entering rows from the grown operand, fill extension from the new V columns) with v, tau
//       read from QT instead of being formed.  Columns are independent under H_k, so the old columns need not be carried: only entries
//       in columns >= n0 are written.
// Requires k0 + l <= n0 (every row a stored reflector touches existed; checked by the caller).
__device__ inline void generic_ab_grow_fix(const AbArgs &a, int l, int u, int r, int n0, int k0, int64_t tile, int lane) {
    constexpr int MB = SSM_MAX_BW, MR = SSM_MAX_RANK;
    const int ub = l + u, nfa = l + u + 1 + r, nfr = nfa, nfq = l + 1;
    const int j0 = n0 - 2 * l - u > 0 ? n0 - 2 * l - u : 0;
    if (k0 <= j0) return;
    const double *au = a.AU + tile * a.np * nfa * 32 + lane;
    const double *vv = a.V + tile * a.np * (r > 0 ? r : 1) * 32 + lane;
    double *ru = a.RU + tile * a.np * nfr * 32 + lane;
    const double *qt = a.QT + tile * a.np * nfq * 32 + lane;
    double W[MB + 1][2 * MB + 1], UC[MB + 1][MR], v[MB];
    for (int i = 0; i <= l; ++i) { for (int t = 0; t <= ub; ++t) W[i][t] = 0.0; for (int q = 0; q < MR; ++q) UC[i][q] = 0.0; }
    // (1) backwards: window rows k .. k+l
    for (int i = 0; i <= l; ++i)
        for (int q = 0; q < r; ++q) UC[i][q] = ru[(((int64_t)k0 - 1 + i) * nfr + ub + 1 + q) * 32];
    for (int k = k0 - 1; k >= j0; --k) {
        const double tau = qt[((int64_t)k * nfq + l) * 32];
        if (tau != 0.0)
            for (int q = 0; q < r; ++q) {
                double s = UC[0][q];
                for (int i = 1; i <= l; ++i) s = fma(qt[((int64_t)k * nfq + i - 1) * 32], UC[i][q], s);
                s *= tau; UC[0][q] -= s;
                for (int i = 1; i <= l; ++i) UC[i][q] = fma(-qt[((int64_t)k * nfq + i - 1) * 32], s, UC[i][q]);
            }
        if (k > j0) {
            for (int i = l; i >= 1; --i) for (int q = 0; q < r; ++q) UC[i][q] = UC[i - 1][q];
            for (int q = 0; q < r; ++q) UC[0][q] = ru[(((int64_t)k - 1) * nfr + ub + 1 + q) * 32];
        }
    }
    // (2) forwards from step j0; at j0 = 0 the window starts from the operand's first rows (the sweep's prologue), otherwise the rows
    // j0 .. j0+l-1 hold nothing in a column >= n0 yet (their last live column is j0 + ub = n0 - l)
    if (j0 == 0)
        for (int rho = 0; rho < l; ++rho)
            for (int t = 0; t <= ub; ++t) {
                if (t - rho <= u) W[rho][t] = t - rho + l >= 0 ? au[((int64_t)rho * nfa + (t - rho + l)) * 32] : 0.0;
                else { double s = 0.0; for (int q = 0; q < r; ++q) s = fma(UC[rho][q], vv[((int64_t)t * r + q) * 32], s); W[rho][t] = s; }
            }
    for (int k = j0; k < k0; ++k) {
        const int64_t e = (int64_t)k + l;
        for (int t = 0; t <= ub; ++t) W[l][t] = (t <= l + u) ? au[(e * nfa + t) * 32] : 0.0;
        for (int q = 0; q < r; ++q) UC[l][q] = au[(e * nfa + l + u + 1 + q) * 32];
        const double tau = qt[((int64_t)k * nfq + l) * 32];
        if (tau != 0.0) {
            for (int i = 1; i <= l; ++i) v[i - 1] = qt[((int64_t)k * nfq + i - 1) * 32];
            for (int t = 1; t <= ub; ++t) {
                double s = W[0][t];
                for (int i = 1; i <= l; ++i) s = fma(v[i - 1], W[i][t], s);
                s *= tau; W[0][t] -= s;
                for (int i = 1; i <= l; ++i) W[i][t] = fma(-v[i - 1], s, W[i][t]);
            }
            for (int q = 0; q < r; ++q) {
                double s = UC[0][q];
                for (int i = 1; i <= l; ++i) s = fma(v[i - 1], UC[i][q], s);
                s *= tau; UC[0][q] -= s;
                for (int i = 1; i <= l; ++i) UC[i][q] = fma(-v[i - 1], s, UC[i][q]);
            }
        }
        for (int t = 1; t <= ub; ++t)
            if (k + t >= n0) ru[((int64_t)k * nfr + t) * 32] = W[0][t];
        const int64_t c = (int64_t)k + 1 + ub;
        for (int i = 1; i <= l; ++i) {
            for (int t = 1; t <= ub; ++t) W[i - 1][t - 1] = W[i][t];
            double s = 0.0;
            for (int q = 0; q < r; ++q) s = fma(UC[i][q], vv[(c * r + q) * 32], s);
            W[i - 1][ub] = s;
            for (int q = 0; q < r; ++q) UC[i - 1][q] = UC[i][q];
        }
    }
    // the partially updated rows k0 .. k0+l-1: window row i = matrix row k0+i, window column t = matrix column k0+t
    for (int i = 0; i < l; ++i)
        for (int t = i; t <= ub; ++t)
            if (k0 + t >= n0) ru[(((int64_t)k0 + i) * nfr + (t - i)) * 32] = W[i][t];
}

// transpose: b <- Q'b (reflectors 0..n-1 in order); else b <- Q b (reverse order).  AlmostBandedMatrix.jl:181,201-207
__device__ inline void generic_ab_qt(const AbSolveArgs &a, int l, bool transpose, int64_t tile, int lane) {
    constexpr int MB = SSM_MAX_BW;
    const int nfq = l + 1, n = a.n;
    const double *qt = a.QT + tile * a.np * nfq * 32 + lane;
    double *rhs = a.rhs + tile * a.np * 32 + lane;
    const double *rin = a.rhs_in + tile * a.np * 32 + lane;
    double RH[MB + 1];
    for (int i = 0; i <= l; ++i) RH[i] = 0.0;
    if (transpose) {
        for (int i = 0; i < l; ++i) RH[i] = rin[(int64_t)i * 32];
        for (int k = 0; k < n; ++k) {
            RH[l] = (k + l < n) ? rin[(int64_t)(k + l) * 32] : 0.0;
            const double tau = qt[((int64_t)k * nfq + l) * 32];
            double s = RH[0];
            for (int i = 1; i <= l; ++i) s = fma(qt[((int64_t)k * nfq + i - 1) * 32], RH[i], s);
            s *= tau;
            rhs[(int64_t)k * 32] = RH[0] - s;
            for (int i = 1; i <= l; ++i) RH[i - 1] = fma(-qt[((int64_t)k * nfq + i - 1) * 32], s, RH[i]);
        }
    } else {
        for (int k = n - 1; k >= 0; --k) {
            for (int i = l; i >= 1; --i) RH[i] = RH[i - 1];
            RH[0] = rin[(int64_t)k * 32];
            const double tau = qt[((int64_t)k * nfq + l) * 32];
            double s = RH[0];
            for (int i = 1; i <= l; ++i) s = fma(qt[((int64_t)k * nfq + i - 1) * 32], RH[i], s);
            s *= tau;
            RH[0] -= s;
            for (int i = 1; i <= l; ++i) RH[i] = fma(-qt[((int64_t)k * nfq + i - 1) * 32], s, RH[i]);
            if (k + l < n) rhs[(int64_t)(k + l) * 32] = RH[l];   // no later reflector touches row k+l
        }
        for (int i = 0; i < l && i < n; ++i) rhs[(int64_t)i * 32] = RH[i];
    }
}

__device__ inline void generic_ab_backsolve(const AbSolveArgs &a, int l, int u, int r, int64_t tile, int lane) {
    constexpr int MB = SSM_MAX_BW, MR = SSM_MAX_RANK;
    const int ub = l + u, nfr = l + u + 1 + r, n = a.n;
    const double *ru = a.RU + tile * a.np * nfr * 32 + lane;
    const double *vv = a.V + tile * a.np * (r > 0 ? r : 1) * 32 + lane;
    double *rhs = a.rhs + tile * a.np * 32 + lane;
    const double *rin = a.rhs_in + tile * a.np * 32 + lane;
    double XW[2 * MB + 1], S[MR];          // XW[t] = x[i+1+t]
    for (int t = 0; t <= ub; ++t) XW[t] = 0.0;
    for (int q = 0; q < MR; ++q) S[q] = 0.0;
    for (int i = n - 1; i >= 0; --i) {
        const int64_t c = (int64_t)i + ub + 1;
        if (c < n) for (int q = 0; q < r; ++q) S[q] = fma(vv[(c * r + q) * 32], XW[ub], S[q]);   // :229-230
        double acc = rin[(int64_t)i * 32];
        for (int q = 0; q < r; ++q) acc = fma(-ru[((int64_t)i * nfr + ub + 1 + q) * 32], S[q], acc);  // :231
        for (int t = ub; t >= 1; --t) acc = fma(-ru[((int64_t)i * nfr + t) * 32], XW[t - 1], acc);    // :233-236
        const double x = a.unit ? acc : acc * (1.0 / ru[((int64_t)i * nfr) * 32]);   // same rounding as bs_step (ab_core.cuh)
        rhs[(int64_t)i * 32] = x;
        for (int t = ub; t >= 1; --t) XW[t] = XW[t - 1];
        XW[0] = x;
    }
}

// (Unit)UpperTriangular(A) \ b and (Unit)LowerTriangular(A) \ b on an UNFACTORED AlmostBandedMatrix
// (AlmostBandedMatrix.jl:242-270).  Lower variants touch the band part only (:258-270).
__device__ inline void generic_ab_tri(const double *AU, const double *V, double *rhs_, int n, int64_t np, int l, int u, int r,
                                      bool upper, bool unit, int64_t tile, int lane) {
    constexpr int MB = SSM_MAX_BW, MR = SSM_MAX_RANK;
    const int nfa = l + u + 1 + r;
    const double *au = AU + tile * np * nfa * 32 + lane;
    const double *vv = V + tile * np * (r > 0 ? r : 1) * 32 + lane;
    double *rhs = rhs_ + tile * np * 32 + lane;
    double XW[MB + 1], S[MR];
    for (int t = 0; t <= MB; ++t) XW[t] = 0.0;
    for (int q = 0; q < MR; ++q) S[q] = 0.0;
    if (upper) {                                  // XW[t] = x[i+1+t], t = 0..u
        for (int i = n - 1; i >= 0; --i) {
            const int64_t c = (int64_t)i + u + 1;
            if (c < n) for (int q = 0; q < r; ++q) S[q] = fma(vv[(c * r + q) * 32], XW[u], S[q]);
            double acc = rhs[(int64_t)i * 32];
            for (int q = 0; q < r; ++q) acc = fma(-au[((int64_t)i * nfa + l + u + 1 + q) * 32], S[q], acc);
            for (int t = u; t >= 1; --t) acc = fma(-au[((int64_t)i * nfa + l + t) * 32], XW[t - 1], acc);
            const double x = unit ? acc : acc * (1.0 / au[((int64_t)i * nfa + l) * 32]);
            rhs[(int64_t)i * 32] = x;
            for (int t = u; t >= 1; --t) XW[t] = XW[t - 1];
            XW[0] = x;
        }
    } else {                                      // XW[t] = x[i-1-t], t = 0..l-1
        for (int i = 0; i < n; ++i) {
            double acc = rhs[(int64_t)i * 32];
            for (int t = l; t >= 1; --t) acc = fma(-au[((int64_t)i * nfa + l - t) * 32], XW[t - 1], acc);
            const double x = unit ? acc : acc * (1.0 / au[((int64_t)i * nfa + l) * 32]);
            rhs[(int64_t)i * 32] = x;
            for (int t = l; t >= 1; --t) XW[t] = XW[t - 1];
            XW[0] = x;
        }
    }
}

}  // namespace ssm
