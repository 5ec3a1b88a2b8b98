// ab_core.cuh — per-system arithmetic of the AlmostBanded hot path, written once and used by every
// kernel variant (direct-load, bulk-async ring) and by the host-side emulation harness in tests/emul/.
//
// Replaces, per system (reference = /root/reference/src/AlmostBandedMatrix.jl):
//   widening copy               :30-38    (implicit: the window materialises fill entries on demand)
//   _almostbanded_qr!           :135-172  (column-at-a-time restatement of the blocked sweep)
//   lmul!(Q', b)                :181,189
//   _almostbanded_upper_ldiv!   :215-240  (row-at-a-time restatement)
//
// Window bookkeeping.  At step k the l+1 active rows are k..k+l and the live columns are k..k+ub
// (ub = l+u).  Row rho lives in slot rho mod (l+1) for its whole life, and entry (rho, c) lives at
// diagonal index dd = c - rho + l of that slot, so NOTHING moves when the window slides: the k loop is
// unrolled by l+1 and every index below is a compile-time constant (registers, no local memory).
//   entering row k+l :  dd = 0..l+u           <- its original band entries A[rho, rho-l .. rho+u]
//   fill extension   :  dd = l+u+1 .. 2l+u    <- Ucur[rho,:] . V[:, c]   (c = k+1+ub at the end of step k)
#pragma once
#include <cmath>
#include "fastmath.cuh"

#if defined(__CUDACC__)
#define SSM_HD __host__ __device__ __forceinline__
#else
#define SSM_HD inline
#endif

namespace ssm {

template <int L, int U, int R>
struct AbDims {
    static constexpr int UB  = L + U;          // upper bandwidth of R
    static constexpr int P   = L + 1;          // active rows = unroll period
    static constexpr int NFA = L + U + 1 + R;  // fields of an input row record: band row + U row
    static constexpr int ND  = 2 * L + U + 1;  // diagonals a row ever holds
    static constexpr int NFR = L + U + 1 + R;  // fields of an R record: R[k,k..k+ub] + Ufinal[k,:]
    static constexpr int NFQ = L + 1;          // fields of a Q record: reflector tail (l) + tau
};

template <int L, int U, int R>
struct AbState {
    using D = AbDims<L, U, R>;
    double A[D::P][D::ND];
    double UC[D::P][R > 0 ? R : 1];
    double RH[D::P];
};

// rows 0..L-1 are resident before step 0; their entries right of the band come from the fill.
// rec[rho][f] : input row record of row rho;  vcol[c][q] : V[q, c] for c = 0..UB (only U+1..UB are read)
template <int L, int U, int R, bool RHS, class RecFn, class VFn, class BFn>
SSM_HD void ab_prologue(AbState<L, U, R> &st, RecFn rec, VFn vcol, BFn bval) {
    using D = AbDims<L, U, R>;
#pragma unroll
    for (int rho = 0; rho < L; ++rho) {
#pragma unroll
        for (int d = 0; d <= L + U; ++d) st.A[rho][d] = rec(rho, d);
#pragma unroll
        for (int q = 0; q < R; ++q) st.UC[rho][q] = rec(rho, L + U + 1 + q);
        st.RH[rho] = RHS ? bval(rho) : 0.0;
#pragma unroll
        for (int c = rho + U + 1; c <= D::UB; ++c) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < R; ++q) s = fma(st.UC[rho][q], vcol(c, q), s);
            st.A[rho][c - rho + L] = s;
        }
    }
}

// One column of the sweep.  PH = k mod (L+1) must be a compile-time constant at the call site (the caller
// unrolls).  in_rec[f]: record of the entering row k+L;  in_v[q] = V[q, k+1+UB];  in_b = b[k+L].
// act = (k < ncols): the reference forms ncols = n-1 reflectors for a real square matrix (:137), tau[n] = 0; a caller of
// _almostbanded_qr!(A, tau, ncols) may ask for fewer (partial factorisation) or for n.
// Outputs: out_r[0..UB] = R[k, k..k+UB], out_r[UB+1+q] = Ufinal[k,q]; out_q[0..L-1] = v tail, out_q[L] = tau;
//          out_c = (Q'b)[k].
template <int L, int U, int R, bool RHS>
SSM_HD void ab_step(AbState<L, U, R> &st, const int PH, const double *in_rec, const double *in_v,
                    const double in_b, const bool act, double *out_r, double *out_q, double &out_c) {
    using D = AbDims<L, U, R>;
    constexpr int P = D::P, UB = D::UB;
    // (1) row k+L enters
    {
        const int es = (PH + L) % P;
#pragma unroll
        for (int d = 0; d <= L + U; ++d) st.A[es][d] = in_rec[d];
#pragma unroll
        for (int q = 0; q < R; ++q) st.UC[es][q] = in_rec[L + U + 1 + q];
        st.RH[es] = RHS ? in_b : 0.0;
    }
    const int s0 = PH % P;
    // (2) reflector from column k: x_i = A[k+i, k]   (LinearAlgebra.reflector!)
    const double x0 = st.A[s0][L];
    double ssq = x0 * x0;
#pragma unroll
    for (int i = 1; i <= L; ++i) { const double xi = st.A[(PH + i) % P][L - i]; ssq = fma(xi, xi, ssq); }
    const bool nz = act && (ssq != 0.0);
    // nu = copysign(|x|, x0), xi = x0 + nu, tau = xi / nu, 1 / xi: branch-free (fastmath.cuh), each within ~1 ulp of the IEEE
    // sqrt / divide sequence.  Measured on B200 (profiles/r1_measurements.md): 15 % faster sweeps when a warp runs alone on its
    // scheduler (<= 2 tiles per SM), no change at the headline size (memory bound)
    const double rs = fast_rsqrt(ssq);
    const double nrm = fast_sqrt_from_rsqrt(ssq, rs);
    const double nu = copysign(nrm, x0);
    const double xi1 = x0 + nu;
    const double inv_xi = nz ? fast_rcp(xi1) : 0.0;
    const double tau = nz ? xi1 * copysign(rs, x0) : 0.0;
    // what the factor storage keeps below the diagonal: v = x_tail / xi, or — for a column without a reflector (k >= ncols of a
    // partial factorisation, or an exactly zero column) — the column itself, as _almostbanded_qr!(A, tau, ncols) leaves it
    const double keep = nz ? inv_xi : 1.0;
    double v[L > 0 ? L : 1], vkeep[L > 0 ? L : 1];
#pragma unroll
    for (int i = 1; i <= L; ++i) { const double xi = st.A[(PH + i) % P][L - i]; v[i - 1] = xi * inv_xi; vkeep[i - 1] = xi * keep; }
    st.A[s0][L] = nz ? -nu : x0;
    // (3) apply H = I - tau v v' to the live columns, to the U rows and to the right-hand side
#pragma unroll
    for (int t = 1; t <= UB; ++t) {
        double s = st.A[s0][t + L];
#pragma unroll
        for (int i = 1; i <= L; ++i) s = fma(v[i - 1], st.A[(PH + i) % P][t - i + L], s);
        s *= tau;
        st.A[s0][t + L] -= s;
#pragma unroll
        for (int i = 1; i <= L; ++i) st.A[(PH + i) % P][t - i + L] = fma(-v[i - 1], s, st.A[(PH + i) % P][t - i + L]);
    }
#pragma unroll
    for (int q = 0; q < R; ++q) {
        double s = st.UC[s0][q];
#pragma unroll
        for (int i = 1; i <= L; ++i) s = fma(v[i - 1], st.UC[(PH + i) % P][q], s);
        s *= tau;
        st.UC[s0][q] -= s;
#pragma unroll
        for (int i = 1; i <= L; ++i) st.UC[(PH + i) % P][q] = fma(-v[i - 1], s, st.UC[(PH + i) % P][q]);
    }
    if (RHS) {
        double s = st.RH[s0];
#pragma unroll
        for (int i = 1; i <= L; ++i) s = fma(v[i - 1], st.RH[(PH + i) % P], s);
        s *= tau;
        st.RH[s0] -= s;
#pragma unroll
        for (int i = 1; i <= L; ++i) st.RH[(PH + i) % P] = fma(-v[i - 1], s, st.RH[(PH + i) % P]);
    }
    // (4) row k is final
#pragma unroll
    for (int t = 0; t <= UB; ++t) out_r[t] = st.A[s0][t + L];
#pragma unroll
    for (int q = 0; q < R; ++q) out_r[UB + 1 + q] = st.UC[s0][q];
#pragma unroll
    for (int i = 1; i <= L; ++i) out_q[i - 1] = vkeep[i - 1];
    out_q[L] = tau;
    out_c = st.RH[s0];
    // (5) surviving rows gain column k+1+UB, which lies in their fill region
#pragma unroll
    for (int i = 1; i <= L; ++i) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < R; ++q) s = fma(st.UC[(PH + i) % P][q], in_v[q], s);
        st.A[(PH + i) % P][UB + 1 - i + L] = s;
    }
}

// ---- b <- Q'b : reflector k touches rows k..k+L of b (banded QRPackedQ' lmul!, :181,189) ----------------
template <int L>
struct QtState { double RH[L + 1]; };

template <int L, class BFn>
SSM_HD void qt_prologue(QtState<L> &st, BFn bval) {
#pragma unroll
    for (int rho = 0; rho < L; ++rho) st.RH[rho] = bval(rho);
}
// in_q[0..L-1] = v tail of column k, in_q[L] = tau_k; in_b = b[k+L]
template <int L>
SSM_HD double qt_step(QtState<L> &st, const int PH, const double *in_q, const double in_b) {
    constexpr int P = L + 1;
    st.RH[(PH + L) % P] = in_b;
    const int s0 = PH % P;
    double s = st.RH[s0];
#pragma unroll
    for (int i = 1; i <= L; ++i) s = fma(in_q[i - 1], st.RH[(PH + i) % P], s);
    s *= in_q[L];
    const double out = st.RH[s0] - s;
#pragma unroll
    for (int i = 1; i <= L; ++i) st.RH[(PH + i) % P] = fma(-in_q[i - 1], s, st.RH[(PH + i) % P]);
    return out;
}

// ---- x <- R \ c, rows n-1 .. 0 (reference blocks of ub+1 rows, :215-240) --------------------
// XW holds x[i+1 .. i+UB+1] in rotating slots: x[j] lives in slot j mod (UB+1).
template <int UB, int R>
struct BsState { double XW[UB + 1]; double S[R > 0 ? R : 1]; };

template <int UB, int R>
SSM_HD void bs_init(BsState<UB, R> &st) {
#pragma unroll
    for (int t = 0; t <= UB; ++t) st.XW[t] = 0.0;
#pragma unroll
    for (int q = 0; q < R; ++q) st.S[q] = 0.0;
}
// Row i, with PH = i mod (UB+1) compile-time.  in_r: R record of row i; in_v[q] = V[q, i+UB+1] (0 beyond n);
// in_c = (Q'b)[i].  unit: UnitUpperTriangular variant (:250-256).
template <int UB, int R>
SSM_HD double bs_step(BsState<UB, R> &st, const int PH, const double *in_r, const double *in_v, const double in_c,
                      const bool unit = false) {
    constexpr int PX = UB + 1;
    // the reciprocal of the diagonal does not depend on the recurrence: formed first, it overlaps the FMA chain below instead of
    // sitting at its end (the back-substitution is a serial chain per system; this matters whenever there are few warps per SM)
    const double rinv = unit ? 1.0 : fast_rcp(in_r[0]);
    // s += V[:, i+UB+1] * x[i+UB+1]; x[i+UB+1] sits in slot (i+UB+1) mod PX = PH  (:229-230)
    const double xfar = st.XW[PH % PX];
#pragma unroll
    for (int q = 0; q < R; ++q) st.S[q] = fma(in_v[q], xfar, st.S[q]);
    double acc = in_c;
#pragma unroll
    for (int q = 0; q < R; ++q) acc = fma(-in_r[UB + 1 + q], st.S[q], acc);   // b[kr] -= U[kr,:]*buffer  (:231)
#pragma unroll
    for (int t = UB; t >= 1; --t) acc = fma(-in_r[t], st.XW[(PH + t) % PX], acc);  // band part (:233-236)
    const double x = acc * rinv;
    st.XW[PH % PX] = x;                                                    // slot of x[i+UB+1] is now x[i]
    return x;
}

}  // namespace ssm
