// bps_core.cuh — per-system arithmetic of the BandedPlusSemiseparable hot path (host+device templates), used by
// the CUDA kernels (bps_kernels.cu) and by the host emulation harness (tests/emul/).
//
// Replaces, per system (reference = /root/reference/src):
//   qr(::BandedPlusSemiseparableMatrix)            semiseparableqr.jl:104-152 (driver), :315-571 (one step),
//                                                  :36-48,:620-640 (A'U columns of S), :573-618 (lookup tables)
//   lmul!(Q', b)                                   semiseparableqr.jl:863-919
//   ldiv!(UpperTriangular{BPS}, b)                 BandedPlusSemiseparableMatrix.jl:53-70
//
// The reference keeps three O(n) lookup tables and an n x r copy of A'U; here every one of them is a running
// quantity so the sweep is a single streaming pass with O(1) state per system:
//   UtU[j+2]  (suffix Gram, :573-585)      = Gtot - sum_{t<=q} u_t u_t'      (Gtot from one reduction pre-pass)
//   uw_sum[q] (prefix, :587-600)           = running sum of u_t w_t'
//   d_extra[q] (:602-618)                  = recomputed from the m-1 rows behind the window (still in the ring)
//   A'U[q,:]  (:42, fast_UtA)              = emitted per step from the same window (rows q-m .. q+l)
// The per-step algebra is the reference's, regrouped: c1 = Q'c3, c2 = K'c3, and the
// w1 / f / c1 / c4 terms enter the state only through phi, the kbar / c2 / c5 terms only through psi.
//
// Row records (all 0 outside the matrix):  field 0..L+M : B0[rho, rho-L+d];  then U[rho,0..R), V[rho,0..R),
// W[rho,0..P), S[rho,0..P).   row(rel, f) returns field f of row q+rel, rel in [-M, L].
#pragma once
#include "fastmath.cuh"
#include <cmath>

#if defined(__CUDACC__)
#define SSM_HD __host__ __device__ __forceinline__
#else
#define SSM_HD inline
#endif

namespace ssm {

template <int L, int M, int R, int P>
struct BpsDims {
    static_assert(L >= 1 && M >= 1 && R >= 1 && P >= 1, "pad degenerate shapes up to (1,1,1,1)");
    static constexpr int LM  = L + M;
    static constexpr int NFB = L + M + 1;                // band fields
    static constexpr int FU = NFB, FV = NFB + R, FW = NFB + 2 * R, FS = NFB + 2 * R + P;
    static constexpr int NFA = NFB + 2 * R + 2 * P;      // input row record
    // output row record, Q part first then R part, so each solve kernel streams exactly the fields it needs:
    //   Q part: reflector tail b(1..L) | Vnew (R) | tau            R part: [pivot, d(1..L+M)] | W alpha (P) | W beta (R) | A'U (R)
    static constexpr int OB = 0, OV = L, OT = L + R, OD = L + R + 1, OWA = OD + L + M + 1, OWB = OWA + P, OSU = OWB + R;
    static constexpr int NFO = OSU + R;
    static constexpr int NFQ = L + R + 1;                // fields of the Q part
    static constexpr int NFRP = NFO - NFQ;               // fields of the R part
};

template <int L, int M, int R, int P>
struct BpsState {
    double Q[R][P], K[R][R], E[R][L + M], X[L][P], Y[L][R], Z[L][L + M];
    double Gpre[R][R];     // sum_{t<=q} u_t u_t'
    double Gtot[R][R];     // sum over all rows (pre-pass)
    double uw[R][P];       // sum_{t<q} u_t w_t'   (original W)
};

template <int L, int M, int R, int P>
SSM_HD void bps_init(BpsState<L, M, R, P> &st, const double *gtot /* R*R row-major */) {
#pragma unroll
    for (int a = 0; a < R; ++a) {
#pragma unroll
        for (int c = 0; c < P; ++c) { st.Q[a][c] = 0.0; st.uw[a][c] = 0.0; }
#pragma unroll
        for (int c = 0; c < R; ++c) { st.K[a][c] = 0.0; st.Gpre[a][c] = 0.0; st.Gtot[a][c] = gtot[a * R + c]; }
#pragma unroll
        for (int t = 0; t < L + M; ++t) st.E[a][t] = 0.0;
    }
#pragma unroll
    for (int s = 0; s < L; ++s) {
#pragma unroll
        for (int c = 0; c < P; ++c) st.X[s][c] = 0.0;
#pragma unroll
        for (int a = 0; a < R; ++a) st.Y[s][a] = 0.0;
#pragma unroll
        for (int t = 0; t < L + M; ++t) st.Z[s][t] = 0.0;
    }
}

// One column q of the sweep.  last = (q == n-1): no reflector, only the final diagonal entry (:123).
// out[] is the output row record (BpsDims::NFO doubles).
//
// PAR = 2: the whole step.  PAR = 0 / 1: this caller owns only the band-wide columns t with (t & 1) == PAR of E, Z, dbar, c6, c3E
// and of the emitted row (out[OD + t], t >= 1) — two warps working on the same systems split the band-wide half of the arithmetic,
// and because a step shifts every column one place left, a warp that owns the odd columns now owns the even ones at the next
// step without moving a register.  Column 0 (needed by the scalar prelude of both) is handed over through `xch`: the caller that
// produces the new column 0 puts it (xch.put(i, v), i < R + L) at the end of its step, the other gets it (xch.get(i)) at the start of
// its next one; the caller must place a barrier between two steps.  Everything that is not band-wide is computed by both callers.
struct BpsNoExchange {
    SSM_HD void put(int, double) const {}
    SSM_HD void sync() const {}
    SSM_HD double get(int) const { return 0.0; }
};
template <int L, int M, int R, int P, int PAR = 2, class RowFn, class XFn = BpsNoExchange>
SSM_HD void bps_step(BpsState<L, M, R, P> &st, RowFn row, const bool last, double *out, XFn xch = XFn()) {
    using D = BpsDims<L, M, R, P>;
    constexpr int LM = D::LM;
#define SSM_OWN(t) (PAR == 2 || (((t) & 1) == PAR))
    // column 0 of E and Z: mine, or handed over by the caller that owns the even columns
    double e0[R], z0[L];
    if (PAR != 1) {
#pragma unroll
        for (int a = 0; a < R; ++a) e0[a] = st.E[a][0];
#pragma unroll
        for (int s = 0; s < L; ++s) z0[s] = st.Z[s][0];
    } else {                       // put by the other caller at the end of its previous step; the caller's barrier lies in between
#pragma unroll
        for (int a = 0; a < R; ++a) e0[a] = xch.get(a);
#pragma unroll
        for (int s = 0; s < L; ++s) z0[s] = xch.get(R + s);
    }
    double u1[R], V1[R], w1[P], S1[P];
#pragma unroll
    for (int a = 0; a < R; ++a) { u1[a] = row(0, D::FU + a); V1[a] = row(0, D::FV + a); }
#pragma unroll
    for (int c = 0; c < P; ++c) { w1[c] = row(0, D::FW + c); S1[c] = row(0, D::FS + c); }
    // suffix Gram G2 = sum_{t>q} u_t u_t'  (:573-585 by subtraction, as fast_UtA does at :627,631)
    double G2[R][R];
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
        for (int c = 0; c < R; ++c) { st.Gpre[a][c] = fma(u1[a], u1[c], st.Gpre[a][c]); G2[a][c] = st.Gtot[a][c] - st.Gpre[a][c]; }
    // rows below the pivot inside the band: Ub[s] = U[q+s,:], xb[s] = B0[q+s, q]
    double Ub[L][R], xb[L];
#pragma unroll
    for (int s = 1; s <= L; ++s) {
        xb[s - 1] = row(s, L - s);
#pragma unroll
        for (int a = 0; a < R; ++a) Ub[s - 1][a] = row(s, D::FU + a);
    }
    // g1 = U[q:q+l,:]' B[q:q+l,q] + G2 V[q,:]   (:326-327)
    double g1[R];
#pragma unroll
    for (int a = 0; a < R; ++a) {
        double s = u1[a] * row(0, L);
#pragma unroll
        for (int t = 1; t <= L; ++t) s = fma(Ub[t - 1][a], xb[t - 1], s);
#pragma unroll
        for (int c = 0; c < R; ++c) s = fma(G2[a][c], V1[c], s);
        g1[a] = s;
    }
    // A'U[q,:] = g1 + U[q-m:q-1,:]' B0[q-m:q-1, q] + (sum_{t<q} u_t w_t') S[q,1:p]     (:42, :633-635)
#pragma unroll
    for (int a = 0; a < R; ++a) {
        double s = g1[a];
#pragma unroll
        for (int t = 1; t <= M; ++t) s = fma(row(-t, D::FU + a), row(-t, L + t), s);
#pragma unroll
        for (int c = 0; c < P; ++c) s = fma(st.uw[a][c], S1[c], s);
        out[D::OSU + a] = s;
    }
    // column q of the trailing block = pivot*e + U^(q+1) kbar + b   (:329-351)
    double core[R], kb[R], b[L];
#pragma unroll
    for (int a = 0; a < R; ++a) {
        double s = e0[a];
#pragma unroll
        for (int c = 0; c < P; ++c) s = fma(st.Q[a][c], S1[c], s);
#pragma unroll
        for (int c = 0; c < R; ++c) s = fma(st.K[a][c], g1[c], s);
        core[a] = s;
        kb[a] = V1[a] + s;
    }
    double pivot = row(0, L) + z0[0];
#pragma unroll
    for (int a = 0; a < R; ++a) { pivot = fma(u1[a], core[a], pivot); pivot = fma(st.Y[0][a], g1[a], pivot); }
#pragma unroll
    for (int c = 0; c < P; ++c) pivot = fma(st.X[0][c], S1[c], pivot);
#pragma unroll
    for (int s = 1; s <= L; ++s) {
        double v = xb[s - 1];
        if (s <= L - 1) {
            v += z0[s];
#pragma unroll
            for (int c = 0; c < P; ++c) v = fma(st.X[s][c], S1[c], v);
#pragma unroll
            for (int a = 0; a < R; ++a) v = fma(st.Y[s][a], g1[a], v);
        }
        b[s - 1] = v;
    }
    if (last) {   // A.B[n,n] = A[n,n]  (:123 via getindex :84-91): rows below do not exist
#pragma unroll
        for (int f = 0; f < D::OSU; ++f) out[f] = 0.0;
        out[D::OD] = pivot;
#pragma unroll
        for (int a = 0; a < R; ++a) out[D::OV + a] = V1[a];       // V[n,:] and W[n,:] are never touched by the sweep
#pragma unroll
        for (int c = 0; c < P; ++c) out[D::OWA + c] = w1[c];
        return;
    }
    // ||column||^2 = pivot^2 + kbar' G2 kbar + 2 kbar' Ub' b + b'b   (:354-357)
    double Ubtb[R], G2k[R];
    double len2 = pivot * pivot;
#pragma unroll
    for (int a = 0; a < R; ++a) {
        double s = 0.0, g = 0.0;
#pragma unroll
        for (int t = 0; t < L; ++t) s = fma(Ub[t][a], b[t], s);
#pragma unroll
        for (int c = 0; c < R; ++c) g = fma(G2[a][c], kb[c], g);
        Ubtb[a] = s; G2k[a] = g;
        len2 = fma(kb[a], g + 2.0 * s, len2);
    }
#pragma unroll
    for (int t = 0; t < L; ++t) len2 = fma(b[t], b[t], len2);
    // reflector (:359-362).  copysign == -sign(pivot) of the reference whenever pivot != 0.
    // Branch-free reciprocal square root / reciprocal (fastmath.cuh, ~1 ulp): sqrt + two IEEE divides are ~220 cycles of this
    // step's serial chain on B200, the seed + Newton forms ~135, and without slow-path branches the compiler overlaps them.
    const double rs = fast_rsqrt(len2);
    const double nu = copysign(fast_sqrt_from_rsqrt(len2, rs), pivot);
    const double den = pivot + nu;
    const double inv = fast_rcp(den);
    const double tau = den * copysign(rs, pivot);
#pragma unroll
    for (int a = 0; a < R; ++a) { kb[a] *= inv; Ubtb[a] *= inv; G2k[a] *= inv; }
#pragma unroll
    for (int t = 0; t < L; ++t) b[t] *= inv;
    // row data (:373-391)
    double d1[M + 1], f[P];
#pragma unroll
    for (int t = 0; t <= M; ++t) d1[t] = row(0, L + t);
#pragma unroll
    for (int c = 0; c < P; ++c) d1[0] = fma(-w1[c], S1[c], d1[0]);
    // suffix sums over the rows below:  PU[t] = sum_{s>t} b_s U[q+s,:],  PW[t] = sum_{s>=max(t,1)} b_s W[q+s,1:p]
    double PU[L][R], PW[L + 1][P];
#pragma unroll
    for (int a = 0; a < R; ++a) PU[L - 1][a] = b[L - 1] * Ub[L - 1][a];
#pragma unroll
    for (int t = L - 2; t >= 0; --t)
#pragma unroll
        for (int a = 0; a < R; ++a) PU[t][a] = fma(b[t], Ub[t][a], PU[t + 1][a]);   // PU[t] = sum_{s>=t+1}, s 1-based
#pragma unroll
    for (int c = 0; c < P; ++c) PW[L][c] = b[L - 1] * row(L, D::FW + c);
#pragma unroll
    for (int t = L - 1; t >= 1; --t)
#pragma unroll
        for (int c = 0; c < P; ++c) PW[t][c] = fma(b[t - 1], row(t, D::FW + c), PW[t + 1][c]);
#pragma unroll
    for (int c = 0; c < P; ++c) { PW[0][c] = PW[1][c]; f[c] = PW[1][c]; }
    // dbar[t] = sum_s b_s A0[q+s, q+t] - S[q+t,1:p] f   (:378-391), t = 0 .. L+M
    double dbar[LM + 1];
#pragma unroll
    for (int t = 0; t <= LM; ++t) {
        if (!SSM_OWN(t)) { dbar[t] = 0.0; continue; }
        double s = 0.0;
#pragma unroll
        for (int i = 1; i <= L; ++i) {
            const int dd = L - i + t;                      // band field of (q+i, q+t)
            if (dd >= 0 && dd <= LM) s = fma(b[i - 1], row(i, dd), s);
        }
        if (t < L) {
#pragma unroll
            for (int a = 0; a < R; ++a) s = fma(PU[t][a], row(t, D::FV + a), s);     // tril part, rows q+s > q+t
        }
        if (t <= L) {
#pragma unroll
            for (int c = 0; c < P; ++c) s = fma(-PW[t][c], row(t, D::FS + c), s);    // triu part cancels for s < t
        }
        dbar[t] = s;
    }
    // y' * pieces (:403-435)
    double c3[R], c1[P], c2[R], c4[P], c5[R], c6[LM], yh[L];
#pragma unroll
    for (int a = 0; a < R; ++a) c3[a] = u1[a] + G2k[a] + Ubtb[a];
#pragma unroll
    for (int c = 0; c < P; ++c) { double s = 0.0;
#pragma unroll
        for (int a = 0; a < R; ++a) s = fma(st.Q[a][c], c3[a], s);
        c1[c] = s; }
#pragma unroll
    for (int c = 0; c < R; ++c) { double s = 0.0;
#pragma unroll
        for (int a = 0; a < R; ++a) s = fma(st.K[a][c], c3[a], s);
        c2[c] = s; }
#pragma unroll
    for (int s = 1; s <= L - 1; ++s) { double v = b[s - 1];
#pragma unroll
        for (int a = 0; a < R; ++a) v = fma(Ub[s - 1][a], kb[a], v);
        yh[s] = v; }
#pragma unroll
    for (int c = 0; c < P; ++c) { double s = st.X[0][c];
#pragma unroll
        for (int t = 1; t <= L - 1; ++t) s = fma(st.X[t][c], yh[t], s);
        c4[c] = s; }
#pragma unroll
    for (int a = 0; a < R; ++a) { double s = st.Y[0][a];
#pragma unroll
        for (int t = 1; t <= L - 1; ++t) s = fma(st.Y[t][a], yh[t], s);
        c5[a] = s; }
#pragma unroll
    for (int t = 0; t < LM; ++t) {
        if (t == 0 || !SSM_OWN(t)) { c6[t] = 0.0; continue; }       // c6[0] is never used
        double s = st.Z[0][t];
#pragma unroll
        for (int i = 1; i <= L - 1; ++i) s = fma(st.Z[i][t], yh[i], s);
        c6[t] = s; }
    double c2u = 0.0, c5u = 0.0, ku = 0.0;
#pragma unroll
    for (int a = 0; a < R; ++a) { c2u = fma(c2[a], u1[a], c2u); c5u = fma(c5[a], u1[a], c5u); ku = fma(kb[a], u1[a], ku); }
    const double sig1 = 1.0 + c2u + c5u;
    double Ku1[R], Yu[L], phi[P], psi[R], c3E[LM];
#pragma unroll
    for (int a = 0; a < R; ++a) { double s = 0.0;
#pragma unroll
        for (int c = 0; c < R; ++c) s = fma(st.K[a][c], u1[c], s);
        Ku1[a] = s; psi[a] = kb[a] + c2[a] + c5[a]; }
#pragma unroll
    for (int s = 1; s <= L - 1; ++s) { double v = 0.0;
#pragma unroll
        for (int a = 0; a < R; ++a) v = fma(st.Y[s][a], u1[a], v);
        Yu[s] = v; }
#pragma unroll
    for (int c = 0; c < P; ++c) phi[c] = fma(sig1, w1[c], f[c] + c1[c] + c4[c]);
#pragma unroll
    for (int t = 0; t < LM; ++t) { double s = 0.0;
        if (t >= 1 && SSM_OWN(t)) {                                   // c3E[0] is never used
#pragma unroll
            for (int a = 0; a < R; ++a) s = fma(c3[a], st.E[a][t], s);
        }
        c3E[t] = s; }
    // row q of R (:518-561) from the pre-update state
    double beta[R];
#pragma unroll
    for (int a = 0; a < R; ++a) { double s = st.Y[0][a];
#pragma unroll
        for (int c = 0; c < R; ++c) s = fma(st.K[c][a], u1[c], s);
        beta[a] = fma(-tau, psi[a], s); }
    const double wcoef = 1.0 - tau + tau * ku;
#pragma unroll
    for (int c = 0; c < P; ++c) { double s = fma(wcoef, w1[c], st.X[0][c]);
#pragma unroll
        for (int a = 0; a < R; ++a) s = fma(st.Q[a][c], u1[a], s);
        s = fma(-tau, f[c] + c1[c] + c4[c], s);
#pragma unroll
        for (int a = 0; a < R; ++a) s = fma(-st.uw[a][c], beta[a], s);
        out[D::OWA + c] = s; }
#pragma unroll
    for (int a = 0; a < R; ++a) out[D::OWB + a] = beta[a];
    out[D::OD] = -nu;
#pragma unroll
    for (int t = 1; t <= LM; ++t) {
        if (!SSM_OWN(t)) continue;
        double s = -tau * dbar[t];
        if (t <= M) s = fma(wcoef, d1[t], s);
        if (t <= LM - 1) {
#pragma unroll
            for (int a = 0; a < R; ++a) s = fma(u1[a], st.E[a][t], s);
            s = fma(-tau, c3E[t] + c6[t], s);
            s += st.Z[0][t];
        }
        if (t <= M) {   // d_extra (:602-618): rows q-(m-1)..q-1, original band entries at column q+t
#pragma unroll
            for (int tp = 1; tp <= M - 1; ++tp) {
                if (tp + t <= M) {
                    double e = 0.0;
#pragma unroll
                    for (int a = 0; a < R; ++a) e = fma(row(-tp, D::FU + a), beta[a], e);
                    s = fma(-e, row(-tp, L + tp + t), s);
                }
            }
        }
        out[D::OD + t] = s;
    }
#pragma unroll
    for (int t = 0; t < L; ++t) out[D::OB + t] = b[t];
#pragma unroll
    for (int a = 0; a < R; ++a) out[D::OV + a] = kb[a];
    out[D::OT] = tau;
    // state recurrences (:441-510)
#pragma unroll
    for (int a = 0; a < R; ++a) {
        const double tk = tau * kb[a];
        const double coefE = fma(-tk, sig1, Ku1[a]);
#pragma unroll
        for (int c = 0; c < P; ++c) st.Q[a][c] = fma(-tk, phi[c], fma(Ku1[a], w1[c], st.Q[a][c]));
#pragma unroll
        for (int c = 0; c < R; ++c) st.K[a][c] = fma(-tk, psi[c], st.K[a][c]);
#pragma unroll
        for (int t = 0; t < LM; ++t) {
            if (!SSM_OWN(t + 1)) continue;
            double v = (t + 1 <= LM - 1) ? st.E[a][t + 1] : 0.0;
            double w = dbar[t + 1];
            if (t + 1 <= LM - 1) w += c3E[t + 1] + c6[t + 1];
            v = fma(-tk, w, v);
            if (t + 1 <= M) v = fma(coefE, d1[t + 1], v);
            st.E[a][t] = v;
        }
    }
#pragma unroll
    for (int s = 1; s <= L; ++s) {        // new row s-1 comes from old row s (one-row shift) plus the rank-one terms of b_s
        const double tb = tau * b[s - 1];
        const bool sh = (s <= L - 1);
#pragma unroll
        for (int c = 0; c < P; ++c) st.X[s - 1][c] = fma(-tb, phi[c], sh ? fma(Yu[s], w1[c], st.X[s][c]) : 0.0);
#pragma unroll
        for (int t = 0; t < LM; ++t) {
            if (!SSM_OWN(t + 1)) continue;
            double v = (sh && t + 1 <= LM - 1) ? st.Z[s][t + 1] : 0.0;
            if (sh && t + 1 <= M) v = fma(Yu[s], d1[t + 1], v);
            double w = dbar[t + 1];
            if (t + 1 <= M) w = fma(sig1, d1[t + 1], w);
            if (t + 1 <= LM - 1) w += c3E[t + 1] + c6[t + 1];
            st.Z[s - 1][t] = fma(-tb, w, v);
        }
#pragma unroll
        for (int a = 0; a < R; ++a) st.Y[s - 1][a] = fma(-tb, psi[a], sh ? st.Y[s][a] : 0.0);
    }
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
        for (int c = 0; c < P; ++c) st.uw[a][c] = fma(u1[a], w1[c], st.uw[a][c]);
    if (PAR == 1) {                // the new column 0 came from old column 1: publish it for the other caller's next step
#pragma unroll
        for (int a = 0; a < R; ++a) xch.put(a, st.E[a][0]);
#pragma unroll
        for (int s = 0; s < L; ++s) xch.put(R + s, st.Z[s][0]);
    }
#undef SSM_OWN
}

// ---- lmul!(Q', b) (semiseparableqr.jl:863-919) as a forward streaming pass ---------------------
// z[s] (s = 0..L) = b[q+s] + pending band corrections; the true vector on rows >= q is z + U h.
template <int L, int R>
struct BpsQtState {
    double z[L + 1];
    double h[R];
    double Gpre[R][R], Gtot[R][R];     // Gram prefix / total of U (as in the QR sweep)
    double tpre[R], ttot[R];           // prefix (rows <= q+L) / total of U' b
};
// before step 0: rows 0..L-1 are in the window.  gtot = U'U (R*R row-major), ttot = U'b (R)
template <int L, int R, class URowFn, class BFn>
SSM_HD void bps_qt_init(BpsQtState<L, R> &st, const double *gtot, const double *ttot, URowFn urow0 /* (row, a) */, BFn b0 /* (row) */) {
#pragma unroll
    for (int a = 0; a < R; ++a) {
        st.h[a] = 0.0; st.tpre[a] = 0.0; st.ttot[a] = ttot[a];
#pragma unroll
        for (int c = 0; c < R; ++c) { st.Gpre[a][c] = 0.0; st.Gtot[a][c] = gtot[a * R + c]; }
    }
#pragma unroll
    for (int s = 0; s < L; ++s) {
        st.z[s] = b0(s);
#pragma unroll
        for (int a = 0; a < R; ++a) st.tpre[a] = fma(urow0(s, a), st.z[s], st.tpre[a]);
    }
    st.z[L] = 0.0;
}
// urow(rel, a) = U[q+rel, a] (rel 0..L);  vq[a] = Vnew[q,a];  bt[s-1] = reflector tail B[q+s, q];  tau;
// b_in = b[q+L] (row entering the window; 0 beyond n).  Returns (Q'b)[q].
template <int L, int R, class URowFn>
SSM_HD double bps_qt_step(BpsQtState<L, R> &st, URowFn urow, const double *vq, const double *bt, const double tau,
                          const double b_in, const bool last) {
    // window slides: z[0..L-1] <- z[1..L] happened at the end of the previous step; row q+L enters with its b
    st.z[L] = b_in;
    double uL[R];
#pragma unroll
    for (int a = 0; a < R; ++a) { uL[a] = urow(L, a); st.tpre[a] = fma(uL[a], b_in, st.tpre[a]); }
    double u0[R];
#pragma unroll
    for (int a = 0; a < R; ++a) u0[a] = urow(0, a);
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
        for (int c = 0; c < R; ++c) st.Gpre[a][c] = fma(u0[a], u0[c], st.Gpre[a][c]);
    double uh0 = 0.0;
#pragma unroll
    for (int a = 0; a < R; ++a) uh0 = fma(u0[a], st.h[a], uh0);
    double out;
    if (last) {
        out = st.z[0] + uh0;                                     // b[n] += U[n,:]'h (:904)
    } else {
        // c = y' x_cur,  y = e_q + tail + U[q+1:,:] vq
        double c = st.z[0] + uh0;
        double w[R];                                             // sum_{rows>q} U[row,:]' x_cur[row]
#pragma unroll
        for (int a = 0; a < R; ++a) {
            double s = st.ttot[a] - st.tpre[a];                  // rows beyond the window: plain b
#pragma unroll
            for (int cc = 0; cc < R; ++cc) s = fma(st.Gtot[a][cc] - st.Gpre[a][cc], st.h[cc], s);   // + G2 h
            w[a] = s;
        }
#pragma unroll
        for (int s = 1; s <= L; ++s) {
            double uh = 0.0;
#pragma unroll
            for (int a = 0; a < R; ++a) { const double us = urow(s, a); uh = fma(us, st.h[a], uh); w[a] = fma(us, st.z[s], w[a]); }
            c = fma(bt[s - 1], st.z[s] + uh, c);
        }
#pragma unroll
        for (int a = 0; a < R; ++a) c = fma(vq[a], w[a], c);
        const double tc = tau * c;
        out = st.z[0] + uh0 - tc;                                // O[q,:] = U[q,:].*h, G[q,1] = -tau c  (:885-890)
#pragma unroll
        for (int a = 0; a < R; ++a) st.h[a] = fma(-tc, vq[a], st.h[a]);          // :887
#pragma unroll
        for (int s = 1; s <= L; ++s) st.z[s] = fma(-tc, bt[s - 1], st.z[s]);     // :892
    }
#pragma unroll
    for (int s = 0; s < L; ++s) st.z[s] = st.z[s + 1];
    return out;
}

// ---- ldiv!(UpperTriangular{BPS}, b) (BandedPlusSemiseparableMatrix.jl:53-70), rows n-1 .. 0 -----------------
// PW = p + r columns of W/S in the factor; MW = upper bandwidth of the factor's B (= L + M for QR factors).
// The serial chain of the recurrence is kept to one FMA and one multiply per row: the W'S coupling with the newest solution
// entry x[j+1] is folded into its band coefficient (sx lags one row behind: sx = sum_{k >= j+2} S[k,:] x[k], sprev = S[j+1,:]),
// everything else of row j is accumulated before x[j+1] is needed, and 1 / B[j,j] is formed off the chain (fastmath.cuh).
template <int MW, int PW>
struct BpsBsState { double xw[MW > 0 ? MW : 1]; double sx[PW]; double sprev[PW]; };   // xw[t] = x[j+1+t]

template <int MW, int PW>
SSM_HD void bps_bs_init(BpsBsState<MW, PW> &st) {
#pragma unroll
    for (int t = 0; t < MW; ++t) st.xw[t] = 0.0;
#pragma unroll
    for (int c = 0; c < PW; ++c) { st.sx[c] = 0.0; st.sprev[c] = 0.0; }
}
// d[0] = B[j,j], d[t] = B[j,j+t] (t = 1..MW); w[c] = W[j,c]; s[c] = S[j,c]; c_in = rhs[j]
template <int MW, int PW>
SSM_HD double bps_bs_step(BpsBsState<MW, PW> &st, const double *d, const double *w, const double *s, const double c_in) {
    const double rinv = fast_rcp(d[0]);                                       // :64, off the chain
    const double xnew = MW > 0 ? st.xw[0] : 0.0;                              // x[j+1], the newest entry
    double coef = MW > 0 ? d[MW > 0 ? 1 : 0] : 0.0;                           // B[j,j+1] + W[j,:]'S[j+1,:]
    double acc = c_in;
#pragma unroll
    for (int c = 0; c < PW; ++c) { acc = fma(-w[c], st.sx[c], acc); coef = fma(w[c], st.sprev[c], coef); }   // W[j,:]'sx  (:60)
#pragma unroll
    for (int t = MW; t >= 2; --t) acc = fma(-d[t], st.xw[t - 1], acc);        // + sum B[j,k] b[k]      (:61-63)
    acc = fma(-coef, xnew, acc);
    const double x = acc * rinv;
#pragma unroll
    for (int c = 0; c < PW; ++c) { st.sx[c] = fma(st.sprev[c], xnew, st.sx[c]); st.sprev[c] = s[c]; }   // sx += S[j+1,:] b[j+1]  (:66)
#pragma unroll
    for (int t = MW - 1; t >= 1; --t) st.xw[t] = st.xw[t - 1];
    if (MW > 0) st.xw[0] = x;
    return x;
}

}  // namespace ssm
