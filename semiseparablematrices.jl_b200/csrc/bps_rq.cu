// bps_rq.cu — the symmetric BandedPlusSemiseparable R*Q step: given the QR factors of a SYMMETRIC
// BandedPlusSemiseparableMatrix, form R*Q = Q'AQ again as a symmetric BandedPlusSemiseparableMatrix — one step of the batched
// QR iteration.  Replaces `rq_mul` (semiseparableqr.jl:282-287) = SymmetricBPSPerturbedRQ constructor (:170-212, fast_RU :818-835),
// `rq_mul!` (:271-280), `onestep_rq!` (:296-311) with compute_gamma (:642-653), compute_variables_delta (:655-685),
// update_lower_triangular_part_rq! (:687-717), update_diagonal_rq! (:719-737), update_next_submatrix_rq! (:739-784), the lookup
// tables (:786-816) and symmetrize_from_lower! (:837-845).
//
// One lane per system, one streaming pass with O(1) state (Omega r x r, Phi l x r, Psi r x l, Lambda l x l): the reference's
// three O(n) tables become "total - running prefix" sums fed by one reduction prologue, exactly as in the QR sweep.  The factor
// is read in place from the device records of ssm_bpsqr (REC: original U / S; OUT: R rows, reflector tails, k-bar, tau, [alpha
// beta], A'U); the result is written straight into the records of a new ssm_bps with bandwidths (l, l) and ranks (r, r)
// (the reference keeps the (l, l+m) storage with zero outer bands; same matrix).  The tests check it against a numpy restatement of the
// reference step and against dense R0*Q.
#include "bps_core.cuh"
#include "ssm_internal.cuh"
#include "stream_loader.cuh"
#include <type_traits>

namespace ssm {

struct RqOut { double *NEW; int64_t np2; int nfa2, tl2, tm2, tr2, tp2; };

// One warp per tile.  The step at row q looks L+M rows ahead into the factor, so the rows q..q+L+M stay in shared memory in
// three bulk-async window rings (stream_loader.cuh WindowRing) that hold exactly the fields the sweep touches:
//   ring A: the R part of the OUT records (pivot/d | alpha | beta | A'U)  — read at rows q..q+L+M
//   ring B: S[:, 1:p] of the REC records                                 — read at rows q..q+L+M
//   ring C: the Q part of the OUT records (reflector tail | k-bar | tau)  — read at row q only (shallow ring)
// Every field is read from HBM once by the sweep (plus the reduction prologue's pass over U, S and A'U); the first version
// read the window rows with direct global loads and spent its time waiting for them (18 ms -> see profiles/r1_measurements.md).
template <int L, int M, int R, int P>
struct RqRings {
    using D = BpsDims<L, M, R, P>;
    static constexpr int MW = L + M;
    static constexpr int NSTA = MW + 6, NSTC = 5;          // window + 5 rows of look-ahead
    using RingA = WindowRing<D::NFRP, NSTA, 1, D::NFO>;
    using RingB = WindowRing<P, NSTA, 1, D::NFA>;
    using RingC = WindowRing<D::NFQ, NSTC, 1, D::NFO>;
    static constexpr size_t OFFB = (RingA::smem_bytes() + 127) / 128 * 128;
    static constexpr size_t OFFC = OFFB + (RingB::smem_bytes() + 127) / 128 * 128;
    static constexpr size_t BYTES = OFFC + RingC::smem_bytes();
};

template <int L, int M, int R, int P>
__global__ void __launch_bounds__(32) bps_rq_kernel(const double *REC, const double *OUT, RqOut o, int *flag, int n, int64_t np, int64_t npo) {
    using D = BpsDims<L, M, R, P>;
    using RR = RqRings<L, M, R, P>;
    static_assert(P == R, "symmetric BPS: p == r");
    constexpr int PW = P + R, MW = L + M;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = (int)threadIdx.x;
    const int64_t tile = blockIdx.x;
    const double *rec_t = REC + (tile * np + BPS_FRONT) * D::NFA * 32;
    const double *out_t = OUT + tile * npo * D::NFO * 32;
    const double *rec0 = rec_t + lane, *out0 = out_t + lane;
    typename RR::RingA ra; typename RR::RingB rb; typename RR::RingC rc;
    ra.init(out_t + D::OD * 32, smem, n);
    rb.init(rec_t + D::FS * 32, smem + RR::OFFB, n);
    rc.init(out_t, smem + RR::OFFC, n);
    // global-memory accessors (prologue)
    auto grec = [&](int i, int f) -> double { return rec0[((int64_t)i * D::NFA + f) * 32]; };
    auto gout = [&](int i, int f) -> double { return out0[((int64_t)i * D::NFO + f) * 32]; };
    constexpr bool writer = true;
    // new operand records
    // The result batch has the padded kernel shape of the factor it comes from (both are bps_pick_shape of a symmetric (l, l, r, r);
    // the launcher checks it), so its record layout is compile-time: no runtime index arithmetic or shape tests around the stores.
    constexpr int FU2 = D::FU, FV2 = D::FV, FW2 = D::FW, FS2 = D::FS;
    double *new0 = o.NEW + (tile * o.np2 + BPS_FRONT) * D::NFA * 32 + lane;
    auto put = [&](int i, int f, double v) { if (writer) new0[((int64_t)i * D::NFA + f) * 32] = v; };

    // ---- prologue: totals for the suffix sums; S[:, 1:r] must equal U (semiseparableqr.jl:199-201) ----
    double T1[R][R], T2[PW][R];
#pragma unroll
    for (int a = 0; a < R; ++a) {
#pragma unroll
        for (int c = 0; c < R; ++c) T1[a][c] = 0.0;
    }
#pragma unroll
    for (int c = 0; c < PW; ++c) {
#pragma unroll
        for (int a = 0; a < R; ++a) T2[c][a] = 0.0;
    }
    bool bad = false;
#pragma unroll 8
    for (int i = 0; i < n; ++i) {
        double su[R], sx[PW];
#pragma unroll
        for (int a = 0; a < R; ++a) { su[a] = grec(i, D::FS + a); bad |= su[a] != grec(i, D::FU + a); sx[a] = su[a]; }
#pragma unroll
        for (int c = P; c < PW; ++c) sx[c] = gout(i, D::OSU + (c - P));
#pragma unroll
        for (int a = 0; a < R; ++a) {
#pragma unroll
            for (int c = 0; c < R; ++c) T1[a][c] = fma(su[a], su[c], T1[a][c]);
        }
#pragma unroll
        for (int c = 0; c < PW; ++c) {
#pragma unroll
            for (int a = 0; a < R; ++a) T2[c][a] = fma(sx[c], su[a], T2[c][a]); }
    }
    if (bad) atomicOr(flag, 1);

    double Om[R][R], Phi[L][R], Psi[R][L], Lam[L][L];
#pragma unroll
    for (int a = 0; a < R; ++a) {
#pragma unroll
        for (int c = 0; c < R; ++c) Om[a][c] = 0.0;
#pragma unroll
        for (int t = 0; t < L; ++t) { Phi[t][a] = 0.0; Psi[a][t] = 0.0; }
    }
#pragma unroll
    for (int s = 0; s < L; ++s) {
#pragma unroll
        for (int t = 0; t < L; ++t) Lam[s][t] = 0.0;
    }

    // One step of the sweep.  EDGE = false: all of rows q..q+MW exist (q + MW < n), so every bounds test of the reference's
    // index ranges folds away at compile time; the last MW steps run the general instance.  Window rows are addressed through
    // per-step slot pointers (row q + t, t a compile-time constant after unrolling): no modulo arithmetic per access.
    int sa = 0, sc = 0;                                   // ring slots of row q (rings A / B, ring C)
    auto step = [&](auto edge, const int q) {
        constexpr bool EDGE = decltype(edge)::value;
        const double *pa[MW + 1], *pb[MW + 1];
#pragma unroll
        for (int t = 0; t <= MW; ++t) {
            int sl = sa + t;
            if (sl >= RR::NSTA) sl -= RR::NSTA;
            pa[t] = ra.ring + sl * (D::NFRP * 32) + lane;
            pb[t] = rb.ring + sl * (P * 32) + lane;
        }
        const double *pc = rc.ring + sc * (D::NFQ * 32) + lane;
        // field f of the OUT record of row i = q + t (rows >= n read as zero); Q-part fields are only read at row q
        auto outf = [&](int i, int f) -> double { return (!EDGE || i < n) ? (f >= D::OD ? pa[i - q][(f - D::OD) * 32] : pc[f * 32]) : 0.0; };
        auto Bup = [&](int i, int k) -> double { return ((!EDGE || k < n) && k - i <= MW) ? outf(i, D::OD + (k - i)) : 0.0; };   // R0[i,k], k >= i
        auto Wext = [&](int i, int c) -> double { return c < P ? outf(i, D::OWA + c) : outf(i, D::OWB + (c - P)); };
        auto Su = [&](int i, int a) -> double { return (!EDGE || i < n) ? pb[i - q][a * 32] : 0.0; };                          // = U0 (checked)
        auto Sext = [&](int i, int c) -> double { return c < P ? Su(i, c) : outf(i, D::OSU + (c - P)); };
        double u1[R];
#pragma unroll
        for (int a = 0; a < R; ++a) u1[a] = Su(q, a);
        // suffix sums over rows > q
#pragma unroll
        for (int a = 0; a < R; ++a) {
#pragma unroll
            for (int c = 0; c < R; ++c) T1[a][c] = fma(-u1[a], u1[c], T1[a][c]);            // UtU[q+1]  (:786-798)
        }
#pragma unroll
        for (int c = 0; c < PW; ++c) { const double s = Sext(q, c);
#pragma unroll
            for (int a = 0; a < R; ++a) T2[c][a] = fma(-s, u1[a], T2[c][a]); }              // S'U over rows > q (:800-816, :818-835)
        const double d0 = Bup(q, q);
        double rho[R];                                                                       // R0[q, q+1:] U0[q+1:, :]
#pragma unroll
        for (int a = 0; a < R; ++a) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < PW; ++c) s = fma(Wext(q, c), T2[c][a], s);
#pragma unroll
            for (int t = 1; t <= MW; ++t) s = fma(Bup(q, q + t), Su(q + t, a), s);
            rho[a] = s;
        }
#pragma unroll
        for (int a = 0; a < R; ++a) { const double un = fma(d0, u1[a], rho[a]); put(q, FU2 + a, un); put(q, FS2 + a, un); }   // U_new = R0 U0 (fast_RU)
        if (EDGE && q == n - 1) {
            // last diagonal entry (:278; getindex :214-269 with k = i = n): the R0 U0 row of the last row is B[n,n] U0[n,:]
            double dg = d0;
#pragma unroll
            for (int a = 0; a < R; ++a) {
                const double ru = d0 * u1[a];
                double s = Psi[a][0];
#pragma unroll
                for (int c = 0; c < R; ++c) s = fma(Om[a][c], u1[c], s);
                dg = fma(ru, s, dg);
                dg = fma(Phi[0][a], u1[a], dg);
            }
            dg += Lam[0][0];
            put(q, L, dg);
#pragma unroll
            for (int a = 0; a < R; ++a) { put(q, FV2 + a, 0.0); put(q, FW2 + a, 0.0); }
            return;
        }
        const int lb = EDGE ? min(L, n - 1 - q) : L, nv = EDGE ? min(L, n - q) : L;
        const double tq = outf(q, D::OT);
        double b[L], kb[R];
#pragma unroll
        for (int t = 0; t < L; ++t) b[t] = t < lb ? outf(q, D::OB + t) : 0.0;
#pragma unroll
        for (int a = 0; a < R; ++a) kb[a] = outf(q, D::OV + a);
        // gamma = R0[q:q+lb, q+1:q+lb] b   (:642-653)
        double gam[L + 1];
#pragma unroll
        for (int t = 0; t <= L; ++t) gam[t] = 0.0;
#pragma unroll
        for (int i = 1; i <= L; ++i) {
            if (i <= lb) {
#pragma unroll
                for (int t = 0; t <= L; ++t) if (t <= i) gam[t] = fma(b[i - 1], Bup(q + t, q + i), gam[t]);
#pragma unroll
                for (int t = 0; t < L; ++t) {
                    if (t < i) {
                        double ws = 0.0;
#pragma unroll
                        for (int c = 0; c < PW; ++c) ws = fma(Wext(q + t, c), Sext(q + i, c), ws);
                        gam[t] = fma(b[i - 1], ws, gam[t]);
                    }
                }
            }
        }
        // delta's (:655-685)
        double d2[R], d1[R], d3[R], d4[L], yh[L];
#pragma unroll
        for (int a = 0; a < R; ++a) {
            double s = u1[a];
#pragma unroll
            for (int c = 0; c < R; ++c) s = fma(T1[a][c], kb[c], s);
#pragma unroll
            for (int t = 0; t < L; ++t) if (t < lb) s = fma(Su(q + 1 + t, a), b[t], s);
            d2[a] = s;
        }
#pragma unroll
        for (int a = 0; a < R; ++a) { double s = 0.0;
#pragma unroll
            for (int c = 0; c < R; ++c) s = fma(Om[a][c], d2[c], s);
            d1[a] = s; }
#pragma unroll
        for (int t = 0; t < L; ++t) {
            double s = 0.0;
            if (t < nv - 1) { s = b[t];
#pragma unroll
                for (int a = 0; a < R; ++a) s = fma(Su(q + 1 + t, a), kb[a], s); }
            yh[t] = s;
        }
#pragma unroll
        for (int a = 0; a < R; ++a) { double s = Psi[a][0];
#pragma unroll
            for (int t = 0; t + 1 < L; ++t) if (t < nv - 1) s = fma(Psi[a][t + 1], yh[t], s);
            d3[a] = s; }
#pragma unroll
        for (int s_ = 0; s_ < L; ++s_) { double s = Lam[s_][0];
#pragma unroll
            for (int t = 0; t + 1 < L; ++t) if (t < nv - 1) s = fma(Lam[s_][t + 1], yh[t], s);
            d4[s_] = s_ < nv ? s : 0.0; }
        // lower triangular part (:687-717)
        double eta[L], mu[R], phid2[L];
        const int k1 = min(lb, L - 1);
#pragma unroll
        for (int t = 0; t < L; ++t) {                 // phid2[t] = Phi[t] . d2
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < R; ++a) s = fma(Phi[t][a], d2[a], s);
            phid2[t] = s;
        }
#pragma unroll
        for (int t = 0; t < L; ++t) {
            double e = -tq * gam[t + 1];
            if (t < k1 && t + 1 < L) {
                double s = 0.0;
#pragma unroll
                for (int a = 0; a < R; ++a) s = fma(Phi[t + 1 < L ? t + 1 : 0][a], u1[a], s);
                e += s - tq * phid2[t + 1 < L ? t + 1 : 0] + Lam[t + 1 < L ? t + 1 : 0][0];
            }
            if (t < nv - 1 && t + 1 < L) e += -tq * d4[t + 1 < L ? t + 1 : 0];
            eta[t] = t < lb ? e : 0.0;
        }
        double kk[R];
#pragma unroll
        for (int a = 0; a < R; ++a) {
            double s = Psi[a][0];
#pragma unroll
            for (int c = 0; c < R; ++c) s = fma(Om[a][c], u1[c], s);
            kk[a] = kb[a] + d1[a] + d3[a];
            mu[a] = s - tq * kk[a];
        }
        // diagonal (:719-737)
        double dg;
        {
            double uOu = 0.0, ud1 = 0.0, ud3 = 0.0, rkb = 0.0, rOu = 0.0, rd1 = 0.0, rd3 = 0.0, p0u = 0.0, uPs = 0.0, rPs = 0.0;
#pragma unroll
            for (int a = 0; a < R; ++a) {
                double Ou = 0.0;
#pragma unroll
                for (int c = 0; c < R; ++c) Ou = fma(Om[a][c], u1[c], Ou);
                uOu = fma(u1[a], Ou, uOu); rOu = fma(rho[a], Ou, rOu);
                ud1 = fma(u1[a], d1[a], ud1); ud3 = fma(u1[a], d3[a], ud3);
                rkb = fma(rho[a], kb[a], rkb); rd1 = fma(rho[a], d1[a], rd1); rd3 = fma(rho[a], d3[a], rd3);
                p0u = fma(Phi[0][a], u1[a], p0u); uPs = fma(u1[a], Psi[a][0], uPs); rPs = fma(rho[a], Psi[a][0], rPs);
            }
            dg = d0 - tq * d0 - tq * rkb + d0 * (uOu - tq * ud1) + (rOu - tq * rd1) - tq * d0 * ud3 - tq * rd3;
            dg += -tq * gam[0] + p0u + d0 * uPs - tq * phid2[0] + rPs - tq * d4[0] + Lam[0][0];
        }
        // write row q of the result: diagonal, V_new = mu (= W_new), and the new sub-diagonal column tail, mirrored
        put(q, L, dg);
#pragma unroll
        for (int a = 0; a < R; ++a) { put(q, FV2 + a, mu[a]); put(q, FW2 + a, mu[a]); }
#pragma unroll
        for (int t = 0; t < L; ++t) {
            if (t < lb) {
                put(q + 1 + t, L - (t + 1), eta[t]);           // A[q+1+t, q]
                put(q, L + (t + 1), eta[t]);                   // A[q, q+1+t]   (symmetrize_from_lower!, :837-845)
            }
        }
        // next submatrix (:739-784)
#pragma unroll
        for (int a = 0; a < R; ++a) {
#pragma unroll
            for (int c = 0; c < R; ++c) Om[a][c] = fma(-tq * kk[a], kb[c], Om[a][c]);
        }
#pragma unroll
        for (int a = 0; a < R; ++a) {
#pragma unroll
            for (int t = 0; t < L; ++t) {
                const double old = t + 1 < L ? Psi[a][t + 1 < L ? t + 1 : 0] : 0.0;
                Psi[a][t] = t < lb ? fma(-tq * kk[a], b[t], old) : 0.0;
            }
        }
        double phid2n[L];                              // Phi_old[s+1] . d2 for the Lambda / Phi updates
#pragma unroll
        for (int s_ = 0; s_ < L; ++s_) phid2n[s_] = s_ + 1 < L ? phid2[s_ + 1 < L ? s_ + 1 : 0] : 0.0;
#pragma unroll
        for (int s_ = 0; s_ < L; ++s_) {
#pragma unroll
            for (int t = 0; t < L; ++t) {
                double v = 0.0;
                if (s_ < lb && t < lb) {
                    v = -tq * gam[s_ + 1] * b[t];
                    if (s_ + 1 < L) v += -tq * phid2n[s_] * b[t];
                    if (s_ + 1 < L && t + 1 < L) v += Lam[s_ + 1 < L ? s_ + 1 : 0][t + 1 < L ? t + 1 : 0];
                    if (s_ < nv - 1 && s_ + 1 < L) v += -tq * d4[s_ + 1 < L ? s_ + 1 : 0] * b[t];
                }
                Lam[s_][t] = v;
            }
        }
        // Phi: rows below lb keep what the reference leaves there (only their first column is cleared, :783)
        double PhiN[L][R];
#pragma unroll
        for (int s_ = 0; s_ < L; ++s_) {
#pragma unroll
            for (int a = 0; a < R; ++a) {
                double v = s_ < lb ? -tq * gam[s_ + 1] * kb[a] : Phi[s_][a];
                if (s_ + 1 < L) v += Phi[s_ + 1 < L ? s_ + 1 : 0][a] - tq * phid2n[s_] * kb[a];
                if (s_ < nv - 1 && s_ + 1 < L) v += -tq * d4[s_ + 1 < L ? s_ + 1 : 0] * kb[a];
                if (s_ >= lb && a == 0) v = 0.0;
                PhiN[s_][a] = v;
            }
        }
#pragma unroll
        for (int s_ = 0; s_ < L; ++s_) {
#pragma unroll
            for (int a = 0; a < R; ++a) Phi[s_][a] = PhiN[s_][a];
        }
    };
    for (int k = 0; k < MW && k < n; ++k) { ra.wait(k); rb.wait(k); }
    for (int q = 0; q < n; ++q) {
        if (q + MW < n) { ra.wait(q + MW); rb.wait(q + MW); }
        rc.wait(q);
        if (q + MW < n) step(std::false_type{}, q);
        else step(std::true_type{}, q);
        // row q leaves the window: re-arm its slots with rows q + NST
        ra.release(q); rb.release(q); rc.release(q);
        if (++sa == RR::NSTA) sa = 0;
        if (++sc == RR::NSTC) sc = 0;
    }
}

// ---- operands -> reference layout (ssm_bps_get) ------------------------------------------------------------------------------
struct BpsShape2 { int n, l, m, r, p, tl, tm, tr, tp; };
__global__ void bps_unpack_kernel(const double *REC, double *B, int64_t sb, double *U, double *V, int64_t suv, double *W, double *S, int64_t sws,
                                  BpsShape2 sh, int64_t nb, int64_t np) {
    const int lane = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y, tile = blockIdx.y, s = tile * TILE + lane;
    if (i >= sh.n || s >= nb) return;
    const int nfa = sh.tl + sh.tm + 1 + 2 * sh.tr + 2 * sh.tp, nfb = sh.tl + sh.tm + 1;
    const double *rec = REC + ((tile * np + BPS_FRONT + i) * nfa) * 32 + lane;
    if (B) for (int d = 0; d < nfb; ++d) {
        const int64_t j = i - sh.tl + d;
        if (j >= 0 && j < sh.n && j - i <= sh.m && i - j <= sh.l) B[s * sb + (sh.m + i - j) + (int64_t)(sh.l + sh.m + 1) * j] = rec[d * 32];
    }
    for (int a = 0; a < sh.r; ++a) {
        if (U) U[s * suv + i + (int64_t)sh.n * a] = rec[(nfb + a) * 32];
        if (V) V[s * suv + i + (int64_t)sh.n * a] = rec[(nfb + sh.tr + a) * 32];
    }
    for (int c = 0; c < sh.p; ++c) {
        if (W) W[s * sws + i + (int64_t)sh.n * c] = rec[(nfb + 2 * sh.tr + c) * 32];
        if (S) S[s * sws + i + (int64_t)sh.n * c] = rec[(nfb + 2 * sh.tr + sh.tp + c) * 32];
    }
}

}  // namespace ssm

using namespace ssm;
#define SSM_CHECK_ARG(cond, msg) do { if (!(cond)) { set_error("%s: %s", __func__, msg); return SSM_E_ARG; } } while (0)

#define SSM_RQ_SHAPES(X) X(1, 1, 1, 1) X(2, 2, 2, 2) X(3, 3, 2, 2) X(4, 5, 3, 3)

extern "C" {

// rq_mul(F.factors, F.tau) (semiseparableqr.jl:282-287): RQ <- R*Q as a new symmetric BandedPlusSemiseparableMatrix batch
// BPS(B (l, l), (U_new, V_new), (V_new, U_new)).  F must be the QR of a symmetric BPS (W = V, S = U, m = l); like the reference
// constructor (:195-205) it refuses factors whose S[:, 1:r] differs from U (SSM_E_STATE) or whose shape is not symmetric.
int ssm_bpsqr_rq_mul(const ssm_bpsqr *F, ssm_bps **out) {
    SSM_CHECK_ARG(F && out, "null argument");
    if (F->p != F->r || F->m != F->l) { set_error("ssm_bpsqr_rq_mul: DimensionMismatch: R*Q needs the factors of a symmetric BandedPlusSemiseparableMatrix (p == r, m == l); got (l,m,r,p)=(%d,%d,%d,%d)", F->l, F->m, F->r, F->p); return SSM_E_SHAPE; }
    if (F->n <= F->tl) { set_error("ssm_bpsqr_rq_mul: n must exceed the lower bandwidth"); return SSM_E_UNSUPPORTED; }
    ssm_ctx *c = F->ctx;
    ssm_bps *Q = *out;                                     // reuse an existing result batch of the same shape (steady-state iteration)
    const bool fresh = Q == nullptr;
    if (Q && (Q->n != F->n || Q->l != F->l || Q->m != F->l || Q->r != F->r || Q->p != F->r || Q->nb != F->nb)) { set_error("ssm_bpsqr_rq_mul: result batch has a different shape"); return SSM_E_SHAPE; }
    if (!Q) SSM_TRY(ssm_bps_create(c, F->n, F->l, F->l, F->r, F->r, F->nb, &Q));
    else SSM_TRY(bps_own_rec(Q));                          // a reused result batch may be the operand of a live factor object (even of F itself)
    Q->gt_valid = false;                                   // its U changes: qr(Q) forms U'U again
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(c->device);
    DevBufPtr flag;
    int rc = dev_alloc(c, 1, true, flag);
    if (!rc) {
        RqOut o{Q->REC->p, Q->np, bps_nfa(Q->tl, Q->tm, Q->tr, Q->tp), Q->tl, Q->tm, Q->tr, Q->tp};
        if (Q->tl != F->tl || Q->tm != F->tm || Q->tr != F->tr || Q->tp != F->tp) {
            set_error("ssm_bpsqr_rq_mul: result batch padded to (%d,%d,%d,%d), factors to (%d,%d,%d,%d)", Q->tl, Q->tm, Q->tr, Q->tp, F->tl, F->tm, F->tr, F->tp);
            cudaSetDevice(prev);
            if (fresh) ssm_bps_destroy(Q);
            return SSM_E_UNSUPPORTED;
        }
        rc = SSM_E_UNSUPPORTED;
#define X(L_, M_, R_, P_)                                                                                                   \
    if (F->tl == L_ && F->tm == M_ && F->tr == R_ && F->tp == P_) {                                                        \
        auto kern = bps_rq_kernel<L_, M_, R_, P_>;                                                                         \
        const size_t smem = RqRings<L_, M_, R_, P_>::BYTES;                                                                \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                                \
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);        \
        kern<<<(unsigned)F->ntiles, 32, smem, c->stream>>>(F->REC->p, F->OUT->p, o, reinterpret_cast<int *>(flag->p), F->n, F->np, F->npo); \
        rc = 0;                                                                                                            \
    }
        SSM_RQ_SHAPES(X)
#undef X
        if (!rc) {
            c->launches++;
            int h = 0;
            cudaError_t e = cudaMemcpyAsync(&h, flag->p, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) rc = cuda_fail(e, "bps_rq_kernel", __FILE__, __LINE__);
            else if (h) { set_error("The first r column of S should be equal to U"); rc = SSM_E_STATE; }      // semiseparableqr.jl:199-201
        }
    }
    cudaSetDevice(prev);
    if (rc) { if (fresh) ssm_bps_destroy(Q); return rc; }
    *out = Q;
    return 0;
}

// operands of a BPS batch back in the reference layout (any pointer may be NULL; host or device)
int ssm_bps_get(const ssm_bps *A, double *B, int64_t sb, double *U, double *V, int64_t suv, double *W, double *S, int64_t sws) {
    SSM_CHECK_ARG(A, "null argument");
    ssm_ctx *c = A->ctx;
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(c->device);
    const size_t eb = (size_t)(A->l + A->m + 1) * A->n, eu = (size_t)A->n * A->r, ew = (size_t)A->n * A->p;
    if (sb == 0) sb = (int64_t)eb;
    if (suv == 0) suv = (int64_t)eu;
    if (sws == 0) sws = (int64_t)ew;
    double *stg = nullptr;
    int rc = staging_dev(c, (size_t)A->nb * (eb + 2 * eu + 2 * ew) + 8, &stg);
    if (!rc) {
        double *sB = stg, *sU = sB + (size_t)A->nb * eb, *sV = sU + (size_t)A->nb * eu, *sW = sV + (size_t)A->nb * eu, *sS = sW + (size_t)A->nb * ew;
        cudaMemsetAsync(sB, 0, (size_t)A->nb * eb * sizeof(double), c->stream);
        BpsShape2 sh{A->n, A->l, A->m, A->r, A->p, A->tl, A->tm, A->tr, A->tp};
        bps_unpack_kernel<<<dim3((unsigned)((A->n + 7) / 8), (unsigned)A->ntiles), dim3(32, 8), 0, c->stream>>>(A->REC->p, B ? sB : nullptr, (int64_t)eb, U ? sU : nullptr,
            V ? sV : nullptr, (int64_t)eu, W ? sW : nullptr, S ? sS : nullptr, (int64_t)ew, sh, A->nb, A->np);
        c->launches++;
        auto outcopy = [&](double *dst, int64_t stride, const double *src, size_t elems) -> cudaError_t {
            if (!dst || elems == 0) return cudaSuccess;
            if ((size_t)stride == elems) return cudaMemcpyAsync(dst, src, elems * A->nb * sizeof(double), cudaMemcpyDefault, c->stream);
            return cudaMemcpy2DAsync(dst, stride * sizeof(double), src, elems * sizeof(double), elems * sizeof(double), A->nb, cudaMemcpyDefault, c->stream);
        };
        cudaError_t e = outcopy(B, sb, sB, eb);
        if (e == cudaSuccess) e = outcopy(U, suv, sU, eu);
        if (e == cudaSuccess) e = outcopy(V, suv, sV, eu);
        if (e == cudaSuccess) e = outcopy(W, sws, sW, ew);
        if (e == cudaSuccess) e = outcopy(S, sws, sS, ew);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "ssm_bps_get", __FILE__, __LINE__);
    }
    cudaSetDevice(prev);
    return rc;
}

}  // extern "C"
