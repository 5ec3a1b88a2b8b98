// bps_kernels.cu — sm_100a kernels of the BandedPlusSemiseparable hot path:
//   K5 bps_reduce   U'U (and U'b) per system: the one reduction the streaming sweep needs up front
//   K6 bps_qr       the per-column state-space sweep (semiseparableqr.jl:104-152,315-571) over a shared-memory window
//                   of row records fed by bulk-async copies (WindowRing, stream_loader.cuh)
//   K7 bps_qt       lmul!(Q', b) (semiseparableqr.jl:863-919) as a forward streaming pass
//   K8 bps_bs       ldiv!(UpperTriangular{BPS}, b) (BandedPlusSemiseparableMatrix.jl:53-70)
// plus the boundary kernels (pack / unpack / generate) and the O(n) matvec / residual (K9).
// One lane per system, one warp per tile of 32 systems, as in the AlmostBanded path.
#include "../../include/ssm_rng.h"
#include <type_traits>
#include "bps_core.cuh"
#include "ssm_internal.cuh"
#include "stream_loader.cuh"

namespace ssm {

// compiled kernel shapes (L, M, R, P); a batch is zero-padded to the smallest one that dominates its shape
#define SSM_BPS_SHAPES(X) X(1, 1, 1, 1) X(2, 2, 2, 2) X(3, 3, 2, 2) X(4, 5, 3, 3)

bool bps_pick_shape(int l, int m, int r, int p, int *tl, int *tm, int *tr, int *tp) {
#define X(L_, M_, R_, P_) \
    if (l <= L_ && m <= M_ && r <= R_ && p <= P_) { *tl = L_; *tm = M_; *tr = R_; *tp = P_; return true; }
    SSM_BPS_SHAPES(X)
#undef X
    return false;
}
int bps_nfa(int tl, int tm, int tr, int tp) { return tl + tm + 1 + 2 * tr + 2 * tp; }
int bps_nfo(int tl, int tm, int tr, int tp) { return tl + tr + 1 + tl + tm + 1 + tp + 2 * tr; }

// ---------------------------------------------------------------------------------------------------------
// K5: per-system reductions over the rows: G = U'U (R x R) and, if rhs != null, t = U'b (R).
// RW warps per tile, each sums a slice of the rows (loads of 8 rows in flight per warp); combined through shared memory.  The pass
// reads 2 of the 15 fields of a record, so it is bound by the latency of strided loads, not by bandwidth: 16 warps per tile
// instead of 4 took it from 0.41 to ~0.15 ms at config 3 (it runs once before the QR sweep and once before every Q'b).
// ---------------------------------------------------------------------------------------------------------
constexpr int BPS_RW = 16;
template <int NFA, int FU, int R>
__global__ void __launch_bounds__(32 * BPS_RW) bps_reduce_kernel(const double *REC, const double *rhs, double *GT, double *TT, int n, int64_t np,
                                                                 int64_t np_vec) {
    __shared__ double sh[BPS_RW][R * R + R][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t tile = blockIdx.x;
    const double *rec = REC + (tile * np + BPS_FRONT) * NFA * 32 + lane;
    const double *bb = rhs ? rhs + tile * np_vec * 32 + lane : nullptr;
    double G[R][R], t[R];
#pragma unroll
    for (int a = 0; a < R; ++a) { t[a] = 0.0;
#pragma unroll
        for (int c = 0; c < R; ++c) G[a][c] = 0.0; }
    const int chunk = (n + BPS_RW - 1) / BPS_RW, i0 = w * chunk, i1 = min(n, i0 + chunk);
#pragma unroll 8
    for (int i = i0; i < i1; ++i) {
        double u[R];
#pragma unroll
        for (int a = 0; a < R; ++a) u[a] = __ldg(rec + ((int64_t)i * NFA + FU + a) * 32);
        if (GT) {
#pragma unroll
            for (int a = 0; a < R; ++a)
#pragma unroll
                for (int c = a; c < R; ++c) G[a][c] = fma(u[a], u[c], G[a][c]);
        }
        if (bb) { const double bi = bb[(int64_t)i * 32];
#pragma unroll
            for (int a = 0; a < R; ++a) t[a] = fma(u[a], bi, t[a]); }
    }
#pragma unroll
    for (int a = 0; a < R; ++a) { sh[w][R * R + a][lane] = t[a];
#pragma unroll
        for (int c = 0; c < R; ++c) sh[w][a * R + c][lane] = (c >= a) ? G[a][c] : G[c][a]; }
    __syncthreads();
    for (int e = w; e < R * R + R; e += BPS_RW) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < BPS_RW; ++k) s += sh[k][e][lane];
        if (e < R * R) { if (GT) GT[(tile * (R * R) + e) * 32 + lane] = s; }
        else if (TT) TT[(tile * R + (e - R * R)) * 32 + lane] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------
// K6: QR sweep
// ---------------------------------------------------------------------------------------------------------
// Window ring depth of the one-warp-per-tile sweep: the M + L + 1 rows a step reads plus ONE slot in flight.  Shared memory per
// CTA decides how many warps an SM holds (16 slots = 61 KB = 3 CTAs, 8 slots = 30.7 KB = 7 CTAs), and with 7 warps to switch
// between deeper prefetch buys nothing: 131,072 systems 41.3 -> 27.7 ms (55 -> 80 % of the HBM peak) on B200.
constexpr int bps_nst(int l, int m, int look) { return l + m + 1 + look; }

// SPLIT = 1: one warp per tile of 32 systems.  SPLIT = 2: a CTA of two warps per tile, each warp owning 16 systems (lanes 16-31
// shadow lanes 0-15 and do not store), both reading one shared window ring: the sweep is latency bound — a batch of 16,384 systems
// is only 512 warps, fewer than the 592 warp schedulers of a B200 — so halving the systems per warp doubles the warps in flight at
// the cost of idle lanes, not of HBM or TMA traffic.
constexpr int BPS_LOOK_SMALL = 6;   // default look-ahead of the one-warp sweep at <= 4 tiles per SM (see qr_shape)
constexpr int BPS_NST_SPLIT = 12;  // 46 KB per CTA: four CTAs (all 512 tiles of config 3) stay co-resident per SM
template <int L, int M, int R, int P, int SPLIT, int LOOK = 1>
__global__ void __launch_bounds__(32 * SPLIT) bps_qr_kernel(const double *REC, const double *GT, double *OUT, int n, int64_t np, int64_t npo) {
    using D = BpsDims<L, M, R, P>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int64_t tile = blockIdx.x;
    const int sl = SPLIT == 1 ? lane : (lane & 15) + 16 * (int)(threadIdx.x >> 5);    // system column of the tile this thread computes
    const bool writer = SPLIT == 1 || lane < 16;
    WindowRing<D::NFA, SPLIT == 1 ? bps_nst(L, M, LOOK) : BPS_NST_SPLIT, SPLIT> win;
    // sequence index k <-> row k - M; record of row rho sits at BPS_FRONT + rho
    win.init(REC + (tile * np + BPS_FRONT - M) * D::NFA * 32, smem, (int64_t)n + L + M);
    win.col = sl;
    double gt[R * R];
#pragma unroll
    for (int e = 0; e < R * R; ++e) gt[e] = __ldg(GT + (tile * (R * R) + e) * 32 + sl);
    BpsState<L, M, R, P> st;
    bps_init<L, M, R, P>(st, gt);
    double *out = OUT + tile * npo * D::NFO * 32 + sl;
    constexpr int NST = SPLIT == 1 ? bps_nst(L, M, LOOK) : BPS_NST_SPLIT;
    for (int k = 0; k < L + M; ++k) win.wait(k);          // rows -M .. L-1
    int s0 = 0, sw = (L + M) % NST;                        // slots of sequence records q (row q-M) and q+L+M (row q+L)
    uint32_t pw = ((L + M) / NST) & 1;
    // one column; LAST is a compile-time flag so that the n-1 ordinary steps carry no test for the final row (a branch in the
    // middle of the step is a scheduling barrier for a warp that runs alone on its scheduler)
    auto step = [&](const int q, auto last_tag) {
        constexpr bool LAST = decltype(last_tag)::value;
        win.wait_slot(sw, pw);                             // row q+L
        const double *rp[L + M + 1];                       // window rows q-M .. q+L
#pragma unroll
        for (int t = 0; t <= L + M; ++t) { int sl = s0 + t; if (sl >= NST) sl -= NST; rp[t] = win.slot_ptr(sl); }
        double o[D::NFO];
        bps_step<L, M, R, P>(st, [&](int rel, int f) { return rp[rel + M][f * 32]; }, LAST, o);
        if (writer) {
#pragma unroll
            for (int f = 0; f < D::NFO; ++f) out[((int64_t)q * D::NFO + f) * 32] = o[f];
        }
        win.release_slot(q, s0);                           // row q-M leaves the window
        if (++s0 == NST) s0 = 0;
        if (++sw == NST) { sw = 0; pw ^= 1u; }
    };
    for (int q = 0; q < n - 1; ++q) step(q, std::false_type{});
    step(n - 1, std::true_type{});
}

// ---------------------------------------------------------------------------------------------------------
// K7: b <- Q'b
// ---------------------------------------------------------------------------------------------------------
template <int L, int M, int R, int P, template <int, int, int, int, int, int, int> class Loader>
__global__ void __launch_bounds__(32) bps_qt_kernel(const double *REC, const double *OUT, const double *GT, const double *TT,
                                                    const double *rhs_in, double *rhs, int n, int64_t np, int64_t npo, int64_t npv, int nstage) {
    using D = BpsDims<L, M, R, P>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const double *urec = REC + ((tile * np + BPS_FRONT) * D::NFA + D::FU) * 32 + lane;   // U fields of row 0
    const double *qrec = OUT + (tile * npo * D::NFO + D::OB) * 32 + lane;                 // Q part of row 0
    const double *rin = rhs_in + tile * npv * 32 + lane;
    double *ro = rhs + tile * npv * 32 + lane;
    double gt[R * R], tt[R];
#pragma unroll
    for (int e = 0; e < R * R; ++e) gt[e] = __ldg(GT + (tile * (R * R) + e) * 32 + lane);
#pragma unroll
    for (int a = 0; a < R; ++a) tt[a] = TT[(tile * R + a) * 32 + lane];
    double uw[L + 1][R];                                   // U rows q .. q+L
#pragma unroll
    for (int s = 0; s < L; ++s)
#pragma unroll
        for (int a = 0; a < R; ++a) uw[s][a] = __ldg(urec + ((int64_t)s * D::NFA + a) * 32);
    BpsQtState<L, R> st;
    bps_qt_init<L, R>(st, gt, tt, [&](int row, int a) { return uw[row][a]; }, [&](int row) { return rin[(int64_t)row * 32]; });
    Loader<R, D::NFQ, 1, +1, D::NFA, D::NFO, 1> ld;
    ld.init(urec + (int64_t)L * D::NFA * 32, qrec, rin + (int64_t)L * 32, smem, nstage, n);
    auto step = [&](const int q, auto last_tag) {          // the final row has its own instance: no test for it in the other n-1 steps
        double un[R], qq[D::NFQ], bb[1];
        ld.acquire(q, un, qq, bb);
#pragma unroll
        for (int a = 0; a < R; ++a) uw[L][a] = un[a];
        const double c = bps_qt_step<L, R>(st, [&](int rel, int a) { return uw[rel][a]; }, qq + D::OV, qq + D::OB, qq[D::OT], bb[0], decltype(last_tag)::value);
        ro[(int64_t)q * 32] = c;
#pragma unroll
        for (int s = 0; s < L; ++s)
#pragma unroll
            for (int a = 0; a < R; ++a) uw[s][a] = uw[s + 1][a];
    };
    for (int q = 0; q < n - 1; ++q) step(q, std::false_type{});
    step(n - 1, std::true_type{});
}

// ---------------------------------------------------------------------------------------------------------
// K8: x <- R \ c with R = triu(B) + triu([alpha beta] [S A'U]', 1)
// ---------------------------------------------------------------------------------------------------------
template <int L, int M, int R, int P, template <int, int, int, int, int, int, int> class Loader>
__global__ void __launch_bounds__(32) bps_bs_kernel(const double *REC, const double *OUT, const double *rhs_in, double *rhs, int n,
                                                    int64_t np, int64_t npo, int64_t npv, int nstage) {
    using D = BpsDims<L, M, R, P>;
    constexpr int LM = L + M, PW = P + R;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const double *srec = REC + ((tile * np + BPS_FRONT + (n - 1)) * D::NFA + D::FS) * 32 + lane;     // S fields of row n-1
    const double *rrec = OUT + ((tile * npo + (n - 1)) * D::NFO + D::OD) * 32 + lane;                // R part of row n-1
    const double *rin = rhs_in + (tile * npv + (n - 1)) * 32 + lane;
    double *ro = rhs + tile * npv * 32 + lane;
    BpsBsState<LM, PW> st;
    bps_bs_init<LM, PW>(st);
    Loader<D::NFRP, P, 1, -1, D::NFO, D::NFA, 1> ld;
    ld.init(rrec, srec, rin, smem, nstage, n);
    for (int g = 0; g < n; ++g) {
        const int j = n - 1 - g;
        double rp[D::NFRP], sp[P], cc[1], sfull[PW];
        ld.acquire(g, rp, sp, cc);
#pragma unroll
        for (int c = 0; c < P; ++c) sfull[c] = sp[c];
#pragma unroll
        for (int a = 0; a < R; ++a) sfull[P + a] = rp[LM + 1 + P + R + a];
        const double x = bps_bs_step<LM, PW>(st, rp, rp + LM + 1, sfull, cc[0]);
        ro[(int64_t)j * 32] = x;
    }
}

// ---------------------------------------------------------------------------------------------------------
// K7 / K8 for small batches: the same steps fed by a GroupRing (stream_loader.cuh) — a producer warp issues the copies, the
// consumer waits once per group of rows instead of once per row.
// ---------------------------------------------------------------------------------------------------------
template <int L, int M, int R, int P>
__global__ void __launch_bounds__(64) bps_qt_group_kernel(const double *REC, const double *OUT, const double *GT, const double *TT,
                                                          const double *rhs_in, double *rhs, int n, int64_t np, int64_t npo, int64_t npv, int nstage, int gm) {
    using D = BpsDims<L, M, R, P>;
    using Ring = GroupRing<R, D::NFQ, 1, +1, 1, D::NFA, D::NFO, 1>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile = blockIdx.x;
    const double *urec_t = REC + ((tile * np + BPS_FRONT) * D::NFA + D::FU) * 32;   // U fields of row 0
    const double *qrec_t = OUT + (tile * npo * D::NFO + D::OB) * 32;                 // Q part of row 0
    const double *rin_t = rhs_in + tile * npv * 32;
    Ring::init_barriers(smem, nstage, gm, 1);
    if (warp == 1) {
        if (lane == 0) Ring::produce(urec_t + (int64_t)L * D::NFA * 32, qrec_t, rin_t + (int64_t)L * 32, smem, nstage, gm, n);
        return;
    }
    const double *urec = urec_t + lane, *rin = rin_t + lane;
    double *ro = rhs + tile * npv * 32 + lane;
    double gt[R * R], tt[R];
#pragma unroll
    for (int e = 0; e < R * R; ++e) gt[e] = __ldg(GT + (tile * (R * R) + e) * 32 + lane);
#pragma unroll
    for (int a = 0; a < R; ++a) tt[a] = TT[(tile * R + a) * 32 + lane];
    double uw[L + 1][R];                                   // U rows q .. q+L
#pragma unroll
    for (int s = 0; s < L; ++s)
#pragma unroll
        for (int a = 0; a < R; ++a) uw[s][a] = __ldg(urec + ((int64_t)s * D::NFA + a) * 32);
    BpsQtState<L, R> st;
    bps_qt_init<L, R>(st, gt, tt, [&](int row, int a) { return uw[row][a]; }, [&](int row) { return rin[(int64_t)row * 32]; });
    Ring ld;
    ld.attach(smem, nstage, gm, n, lane);
    auto step = [&](const int q, auto last_tag) {
        double un[R], qq[D::NFQ], bb[1];
        ld.block_begin();
        ld.read(0, un, qq, bb);
        ld.block_end();
#pragma unroll
        for (int a = 0; a < R; ++a) uw[L][a] = un[a];
        const double c = bps_qt_step<L, R>(st, [&](int rel, int a) { return uw[rel][a]; }, qq + D::OV, qq + D::OB, qq[D::OT], bb[0], decltype(last_tag)::value);
        ro[(int64_t)q * 32] = c;
#pragma unroll
        for (int s = 0; s < L; ++s)
#pragma unroll
            for (int a = 0; a < R; ++a) uw[s][a] = uw[s + 1][a];
    };
    for (int q = 0; q < n - 1; ++q) step(q, std::false_type{});
    step(n - 1, std::true_type{});
}

template <int L, int M, int R, int P>
__global__ void __launch_bounds__(64) bps_bs_group_kernel(const double *REC, const double *OUT, const double *rhs_in, double *rhs, int n,
                                                          int64_t np, int64_t npo, int64_t npv, int nstage, int gm) {
    using D = BpsDims<L, M, R, P>;
    constexpr int LM = L + M, PW = P + R;
    using Ring = GroupRing<D::NFRP, P, 1, -1, 1, D::NFO, D::NFA, 1>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile = blockIdx.x;
    const double *srec_t = REC + ((tile * np + BPS_FRONT + (n - 1)) * D::NFA + D::FS) * 32;     // S fields of row n-1
    const double *rrec_t = OUT + ((tile * npo + (n - 1)) * D::NFO + D::OD) * 32;                // R part of row n-1
    const double *rin_t = rhs_in + (tile * npv + (n - 1)) * 32;
    Ring::init_barriers(smem, nstage, gm, 1);
    if (warp == 1) {
        if (lane == 0) Ring::produce(rrec_t, srec_t, rin_t, smem, nstage, gm, n);
        return;
    }
    double *ro = rhs + tile * npv * 32 + lane;
    BpsBsState<LM, PW> st;
    bps_bs_init<LM, PW>(st);
    Ring ld;
    ld.attach(smem, nstage, gm, n, lane);
    for (int g = 0; g < n; ++g) {
        const int j = n - 1 - g;
        double rp[D::NFRP], sp[P], cc[1], sfull[PW];
        ld.block_begin();
        ld.read(0, rp, sp, cc);
        ld.block_end();
#pragma unroll
        for (int c = 0; c < P; ++c) sfull[c] = sp[c];
#pragma unroll
        for (int a = 0; a < R; ++a) sfull[P + a] = rp[LM + 1 + P + R + a];
        const double x = bps_bs_step<LM, PW>(st, rp, rp + LM + 1, sfull, cc[0]);
        ro[(int64_t)j * 32] = x;
    }
}

// ---------------------------------------------------------------------------------------------------------
// boundary kernels (runtime dims = the padded kernel shape): thread = (tile, row, lane)
// ---------------------------------------------------------------------------------------------------------
struct BpsShape { int n, l, m, r, p, tl, tm, tr, tp; };

__global__ void bps_pack_kernel(double *REC, const double *B, int64_t sb, const double *U, const double *V, int64_t suv, const double *W,
                                const double *S, int64_t sws, BpsShape sh, int64_t nb, int64_t np) {
    const int lane = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y, tile = blockIdx.y, s = tile * TILE + lane;
    if (i >= sh.n) return;
    const int nfa = sh.tl + sh.tm + 1 + 2 * sh.tr + 2 * sh.tp, nfb = sh.tl + sh.tm + 1;
    double *rec = REC + ((tile * np + BPS_FRONT + i) * nfa) * 32 + lane;
    const bool live = s < nb;
    for (int d = 0; d < nfb; ++d) {
        const int64_t j = i - sh.tl + d;
        double v = 0.0;
        if (live) { if (j >= 0 && j < sh.n && j - i <= sh.m && i - j <= sh.l) v = B[s * sb + (sh.m + i - j) + (int64_t)(sh.l + sh.m + 1) * j]; }
        else v = (d == sh.tl) ? 1.0 : 0.0;       // padding lanes: identity systems
        rec[d * 32] = v;
    }
    for (int a = 0; a < sh.tr; ++a) {
        rec[(nfb + a) * 32] = (live && a < sh.r) ? U[s * suv + i + (int64_t)sh.n * a] : 0.0;
        rec[(nfb + sh.tr + a) * 32] = (live && a < sh.r) ? V[s * suv + i + (int64_t)sh.n * a] : 0.0;
    }
    for (int c = 0; c < sh.tp; ++c) {
        rec[(nfb + 2 * sh.tr + c) * 32] = (live && c < sh.p) ? W[s * sws + i + (int64_t)sh.n * c] : 0.0;
        rec[(nfb + 2 * sh.tr + sh.tp + c) * 32] = (live && c < sh.p) ? S[s * sws + i + (int64_t)sh.n * c] : 0.0;
    }
}

__global__ void bps_generate_kernel(double *REC, BpsShape sh, int64_t nb, int64_t np, uint64_t seed) {
    const int lane = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y, tile = blockIdx.y, s = tile * TILE + lane;
    if (i >= sh.n) return;
    const int nfa = sh.tl + sh.tm + 1 + 2 * sh.tr + 2 * sh.tp, nfb = sh.tl + sh.tm + 1;
    double *rec = REC + ((tile * np + BPS_FRONT + i) * nfa) * 32 + lane;
    const bool live = s < nb;
    const uint64_t key = ssm_syskey(seed, (uint64_t)s);
    const double su = 1.0 / sqrt((double)(sh.r > 0 ? sh.r : 1) * (double)sh.n), sw = 1.0 / sqrt((double)(sh.p > 0 ? sh.p : 1) * (double)sh.n);
    for (int d = 0; d < nfb; ++d) {
        const int64_t j = i - sh.tl + d;
        double v = (d == sh.tl) ? 1.0 : 0.0;
        if (live) v = (j >= 0 && j < sh.n && j - i <= sh.m && i - j <= sh.l) ? ssm_bps_band_elem(key, sh.n, sh.l, sh.m, (int)(sh.m + i - j), (int)j) : 0.0;
        rec[d * 32] = v;
    }
    for (int a = 0; a < sh.tr; ++a) {
        rec[(nfb + a) * 32] = (live && a < sh.r) ? ssm_bps_gen_elem(key, 1, sh.n, sh.r, (int)i, a, su) : 0.0;
        rec[(nfb + sh.tr + a) * 32] = (live && a < sh.r) ? ssm_bps_gen_elem(key, 2, sh.n, sh.r, (int)i, a, su) : 0.0;
    }
    for (int c = 0; c < sh.tp; ++c) {
        rec[(nfb + 2 * sh.tr + c) * 32] = (live && c < sh.p) ? ssm_bps_gen_elem(key, 3, sh.n, sh.p, (int)i, c, sw) : 0.0;
        rec[(nfb + 2 * sh.tr + sh.tp + c) * 32] = (live && c < sh.p) ? ssm_bps_gen_elem(key, 4, sh.n, sh.p, (int)i, c, sw) : 0.0;
    }
}

// factors -> reference layout of the ORIGINAL shape: B (l, l+m) band storage (2l+m+1) x n; U,V n x r; W,S n x (p+r); tau n
__global__ void bpsqr_unpack_kernel(const double *REC, const double *OUT, double *B, int64_t sb, double *U, double *V, int64_t suv, double *W,
                                    double *S, int64_t sws, double *tau, int64_t st, BpsShape sh, int64_t nb, int64_t np, int64_t npo) {
    const int lane = threadIdx.x;
    const int64_t q = (int64_t)blockIdx.x * blockDim.y + threadIdx.y, tile = blockIdx.y, s = tile * TILE + lane;
    if (q >= sh.n || s >= nb) return;
    const int nfa = sh.tl + sh.tm + 1 + 2 * sh.tr + 2 * sh.tp, nfb = sh.tl + sh.tm + 1;
    const int nfo = sh.tl + sh.tr + 1 + sh.tl + sh.tm + 1 + sh.tp + 2 * sh.tr;
    const int OB = 0, OV = sh.tl, OT = sh.tl + sh.tr, OD = OT + 1, OWA = OD + sh.tl + sh.tm + 1, OWB = OWA + sh.tp, OSU = OWB + sh.tr;
    const double *rec = REC + ((tile * np + BPS_FRONT + q) * nfa) * 32 + lane;
    const double *o = OUT + ((tile * npo + q) * nfo) * 32 + lane;
    const int mw = sh.l + sh.m, nd = 2 * sh.l + sh.m + 1, n = sh.n;
    if (B) {
        for (int t = 0; t <= mw; ++t) if (q + t < n) B[s * sb + (mw - t) + (int64_t)nd * (q + t)] = o[(OD + t) * 32];
        for (int i = 1; i <= sh.l; ++i) if (q + i < n) B[s * sb + (mw + i) + (int64_t)nd * q] = o[(OB + i - 1) * 32];
    }
    for (int a = 0; a < sh.r; ++a) {
        if (U) U[s * suv + q + (int64_t)n * a] = rec[(nfb + a) * 32];
        if (V) V[s * suv + q + (int64_t)n * a] = o[(OV + a) * 32];
        if (W) W[s * sws + q + (int64_t)n * (sh.p + a)] = o[(OWB + a) * 32];
        if (S) S[s * sws + q + (int64_t)n * (sh.p + a)] = o[(OSU + a) * 32];
    }
    for (int c = 0; c < sh.p; ++c) {
        if (W) W[s * sws + q + (int64_t)n * c] = o[(OWA + c) * 32];
        if (S) S[s * sws + q + (int64_t)n * c] = rec[(nfb + 2 * sh.tr + sh.tp + c) * 32];
    }
    if (tau) tau[s * st + q] = o[OT * 32];
}

// K9: y = A x (BandedPlusSemiseparableMatrix.jl:21-41); optionally relres = ||b - A x|| / ||b||
__global__ void __launch_bounds__(32) bps_mul_kernel(const double *REC, const double *x_, double *y_, const double *b_, double *relres, BpsShape sh,
                                                     int64_t nb, int64_t np, int64_t npv) {
    const int lane = threadIdx.x;
    const int64_t tile = blockIdx.x, s = tile * TILE + lane;
    const int nfa = sh.tl + sh.tm + 1 + 2 * sh.tr + 2 * sh.tp, nfb = sh.tl + sh.tm + 1, n = sh.n;
    const double *rec = REC + (tile * np + BPS_FRONT) * nfa * 32 + lane;
    const double *x = x_ + tile * npv * 32 + lane;
    double Stx[SSM_MAX_RANK], Vtx[SSM_MAX_RANK];
    for (int c = 0; c < SSM_MAX_RANK; ++c) { Stx[c] = 0.0; Vtx[c] = 0.0; }
    for (int i = 0; i < n; ++i) { const double xi = x[(int64_t)i * 32];
        for (int c = 0; c < sh.tp; ++c) Stx[c] = fma(rec[((int64_t)i * nfa + nfb + 2 * sh.tr + sh.tp + c) * 32], xi, Stx[c]); }
    double num = 0.0, den = 0.0;
    for (int k = 0; k < n; ++k) {
        const double *rk = rec + (int64_t)k * nfa * 32;
        const double xk = x[(int64_t)k * 32];
        double acc = 0.0;
        for (int d = 0; d < nfb; ++d) { const int64_t j = k - sh.tl + d; if (j >= 0 && j < n) acc = fma(rk[d * 32], x[j * 32], acc); }
        for (int c = 0; c < sh.tp; ++c) Stx[c] = fma(-xk, rk[(nfb + 2 * sh.tr + sh.tp + c) * 32], Stx[c]);
        for (int a = 0; a < sh.tr; ++a) acc = fma(rk[(nfb + a) * 32], Vtx[a], acc);
        for (int c = 0; c < sh.tp; ++c) acc = fma(rk[(nfb + 2 * sh.tr + c) * 32], Stx[c], acc);
        for (int a = 0; a < sh.tr; ++a) Vtx[a] = fma(xk, rk[(nfb + sh.tr + a) * 32], Vtx[a]);
        if (y_) y_[(tile * npv + k) * 32 + lane] = acc;
        if (b_) { const double bk = b_[(tile * npv + k) * 32 + lane], dl = bk - acc; num = fma(dl, dl, num); den = fma(bk, bk, den); }
    }
    if (relres && s < nb) relres[s] = sqrt(num) / sqrt(den);
}

// ldiv!(UpperTriangular(A), b) on an unfactored BPS (test_bandedplussemiseparable.jl:17)
__global__ void __launch_bounds__(32) bps_upper_generic_kernel(const double *REC, double *rhs_, BpsShape sh, int64_t np, int64_t npv) {
    const int lane = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int nfa = sh.tl + sh.tm + 1 + 2 * sh.tr + 2 * sh.tp, nfb = sh.tl + sh.tm + 1, n = sh.n;
    const double *rec = REC + (tile * np + BPS_FRONT) * nfa * 32 + lane;
    double *rhs = rhs_ + tile * npv * 32 + lane;
    double sx[SSM_MAX_RANK];
    for (int c = 0; c < SSM_MAX_RANK; ++c) sx[c] = 0.0;
    for (int j = n - 1; j >= 0; --j) {
        const double *rj = rec + (int64_t)j * nfa * 32;
        double acc = rhs[(int64_t)j * 32];
        for (int c = 0; c < sh.tp; ++c) acc = fma(-rj[(nfb + 2 * sh.tr + c) * 32], sx[c], acc);
        for (int t = 1; t <= sh.tm && j + t < n; ++t) acc = fma(-rj[(sh.tl + t) * 32], rhs[(int64_t)(j + t) * 32], acc);
        const double xj = acc / rj[sh.tl * 32];
        rhs[(int64_t)j * 32] = xj;
        for (int c = 0; c < sh.tp; ++c) sx[c] = fma(rj[(nfb + 2 * sh.tr + sh.tp + c) * 32], xj, sx[c]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------
static BpsShape shape_of(const ssm_bps *A) { return BpsShape{A->n, A->l, A->m, A->r, A->p, A->tl, A->tm, A->tr, A->tp}; }
static BpsShape shape_of(const ssm_bpsqr *F) { return BpsShape{F->n, F->l, F->m, F->r, F->p, F->tl, F->tm, F->tr, F->tp}; }
static dim3 rows_grid(int n, int64_t ntiles) { return dim3((unsigned)((n + 7) / 8), (unsigned)ntiles); }

int launch_bps_pack(ssm_ctx *c, ssm_bps *A, const double *B, int64_t sb, const double *U, const double *V, int64_t suv, const double *W,
                    const double *S, int64_t sws) {
    bps_pack_kernel<<<rows_grid(A->n, A->ntiles), dim3(32, 8), 0, c->stream>>>(A->REC->p, B, sb, U, V, suv, W, S, sws, shape_of(A), A->nb, A->np);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}
int launch_bps_generate(ssm_ctx *c, ssm_bps *A, uint64_t seed) {
    bps_generate_kernel<<<rows_grid(A->n, A->ntiles), dim3(32, 8), 0, c->stream>>>(A->REC->p, shape_of(A), A->nb, A->np, seed);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}
int launch_bpsqr_unpack(ssm_ctx *c, const ssm_bpsqr *F, double *B, int64_t sb, double *U, double *V, int64_t suv, double *W, double *S,
                        int64_t sws, double *tau, int64_t st) {
    bpsqr_unpack_kernel<<<rows_grid(F->n, F->ntiles), dim3(32, 8), 0, c->stream>>>(F->REC->p, F->OUT->p, B, sb, U, V, suv, W, S, sws, tau, st,
                                                                                    shape_of(F), F->nb, F->np, F->npo);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}

// Pair mode: a CTA of two warps sweeps ONE tile together — both warps hold all 32 systems, share the window ring, compute the scalar
// part of every step redundantly and split the band-wide half of the arithmetic (the E / Z state, dbar, c6, the emitted R row) by
// column parity (bps_core.cuh, PAR).  Twice the warps in flight like the 16+16 split above, but no idle lanes in the band-wide part.
// The only traffic between the two warps is column 0 of E and Z (R + L doubles per system and step) through shared memory.
struct BpsPairExchange {
    double *buf; int lane;
    __device__ __forceinline__ void put(int i, double v) const { buf[i * 32 + lane] = v; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ double get(int i) const { return buf[i * 32 + lane]; }
};
template <int L, int M, int R, int P>
__global__ void __launch_bounds__(64) bps_qr_pair_kernel(const double *REC, const double *GT, double *OUT, int n, int64_t np, int64_t npo) {
    using D = BpsDims<L, M, R, P>;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ double xbuf[(R + L) * 32];
    const int lane = threadIdx.x & 31, h = threadIdx.x >> 5;
    const int64_t tile = blockIdx.x;
    WindowRing<D::NFA, BPS_NST_SPLIT, 2> win;
    win.init(REC + (tile * np + BPS_FRONT - M) * D::NFA * 32, smem, (int64_t)n + L + M);
    double gt[R * R];
#pragma unroll
    for (int e = 0; e < R * R; ++e) gt[e] = __ldg(GT + (tile * (R * R) + e) * 32 + lane);
    BpsState<L, M, R, P> st;
    bps_init<L, M, R, P>(st, gt);
    double *out = OUT + tile * npo * D::NFO * 32 + lane;
    const BpsPairExchange xch{xbuf, lane};
    for (int i = threadIdx.x; i < (R + L) * 32; i += 64) xbuf[i] = 0.0;       // column 0 of the zero initial state
    __syncthreads();
    for (int k = 0; k < L + M; ++k) win.wait(k);          // rows -M .. L-1
    // one column; PARC::value = parity of the band-wide columns this warp owns at this step = (h + q) & 1
    constexpr int NST = BPS_NST_SPLIT;
    int s0 = 0, sw = (L + M) % NST;                        // slots of sequence records q (row q-M) and q+L+M (row q+L)
    uint32_t pw = ((L + M) / NST) & 1;
    auto step = [&](auto parc, const int q, auto last_tag) {
        constexpr int PAR = decltype(parc)::value;
        constexpr bool LAST = decltype(last_tag)::value;
        win.wait_slot(sw, pw);                             // row q+L
        const double *rp[L + M + 1];                       // window rows q-M .. q+L
#pragma unroll
        for (int t = 0; t <= L + M; ++t) { int sl = s0 + t; if (sl >= NST) sl -= NST; rp[t] = win.slot_ptr(sl); }
        double o[D::NFO];
        bps_step<L, M, R, P, PAR>(st, [&](int rel, int f) { return rp[rel + M][f * 32]; }, LAST, o, xch);
#pragma unroll
        for (int f = 0; f < D::NFO; ++f) {
            const bool band = f > D::OD && f <= D::OD + L + M;                       // R[q, q+t], t >= 1: written by the column's owner
            const bool mine = band ? (((f - D::OD) & 1) == PAR) : (h == 0);
            if (mine) out[((int64_t)q * D::NFO + f) * 32] = o[f];
        }
        win.release_slot(q, s0);                           // row q-M leaves the window
        if (++s0 == NST) s0 = 0;
        if (++sw == NST) { sw = 0; pw ^= 1u; }
    };
    // the last column runs the LAST instance; all others carry no test for it
    int q = 0;
    if (h == 1) { if (n > 1) step(std::integral_constant<int, 1>{}, 0, std::false_type{}); else step(std::integral_constant<int, 1>{}, 0, std::true_type{}); q = 1; }
    while (q + 2 < n) {
        step(std::integral_constant<int, 0>{}, q++, std::false_type{});
        step(std::integral_constant<int, 1>{}, q++, std::false_type{});
    }
    // 1 or 2 columns left (q = n-2 or n-1); parity of the band-wide columns alternates: PAR = 0 at this q by construction
    if (q + 1 < n) { step(std::integral_constant<int, 0>{}, q++, std::false_type{}); step(std::integral_constant<int, 1>{}, q++, std::true_type{}); }
    else if (q < n) step(std::integral_constant<int, 0>{}, q++, std::true_type{});
}

template <int L, int M, int R, int P>
static int qr_shape(ssm_ctx *c, const ssm_bps *A, ssm_bpsqr *F) {
    using D = BpsDims<L, M, R, P>;
    if (!A->gt_valid) SSM_TRY(launch_bps_gram(c, A));      // operands written by another kernel (rq_mul): form U'U now
    F->GT = A->GT;
    // ab_variant: 0 auto, 1 one warp per tile, 2 two warps of 16 systems, 3 pair mode.  Measured on B200 (n = 4096, (3,3,2,2)):
    //   systems    4,096   8,192   12,288   16,384 (config 3)   32,768   65,536
    //   1 warp     4.28    4.29    4.33     4.45                6.65     13.7  ms
    //   pair       3.81    3.84    4.98     4.99                9.77     19.7  ms
    // A warp alone on its scheduler is bound by the step's serial chain (pair mode shortens it); from two warps per scheduler on,
    // occupancy wins — the 8-slot window ring (30.7 KB) keeps 7 single-warp CTAs per SM resident, pair CTAs are register-limited
    // to 4.  So pair mode only while its two warps per tile still find a scheduler each: at most 2 tiles per SM.
    const bool split = c->ab_variant == 2;
    const bool pair = c->ab_variant == 3 || (c->ab_variant == 0 && A->ntiles <= (int64_t)c->sm_count * 2);
    if (pair) {
        const size_t smem = WindowRing<D::NFA, BPS_NST_SPLIT, 2>::smem_bytes();
        auto kern = bps_qr_pair_kernel<L, M, R, P>;
        SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)A->ntiles, 64, smem, c->stream>>>(A->REC->p, F->GT->p, F->OUT->p, A->n, A->np, F->npo);
    } else if (split) {
        const size_t smem = WindowRing<D::NFA, BPS_NST_SPLIT, 2>::smem_bytes();
        auto kern = bps_qr_kernel<L, M, R, P, 2>;
        SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)A->ntiles, 64, smem, c->stream>>>(A->REC->p, F->GT->p, F->OUT->p, A->n, A->np, F->npo);
    } else {
        // one slot in flight keeps 7 CTAs per SM resident; a second one (6 CTAs per SM) removes most of the wait for the entering
        // row when the batch leaves room for it (config 3: 4.20 -> 3.87 ms; at 7 tiles per SM it would cost a second wave)
        // pipeline_stages > 0 picks the look-ahead by hand (1, 2, 4 or 6 slots in flight): with one warp per scheduler the bytes in flight per SM
        // are tiles/SM * look-ahead * 3.8 KB, and Little's law wants ~6.5 MB device-wide for the HBM rate
        const int look = c->pipeline_stages > 0 ? c->pipeline_stages : (A->ntiles <= (int64_t)c->sm_count * 4 ? BPS_LOOK_SMALL : (A->ntiles <= (int64_t)c->sm_count * 5 ? 2 : 1));
        if (look >= 6) {
            const size_t smem = WindowRing<D::NFA, bps_nst(L, M, 6), 1>::smem_bytes();
            auto kern = bps_qr_kernel<L, M, R, P, 1, 6>;
            SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)A->ntiles, 32, smem, c->stream>>>(A->REC->p, F->GT->p, F->OUT->p, A->n, A->np, F->npo);
        } else if (look >= 4) {
            const size_t smem = WindowRing<D::NFA, bps_nst(L, M, 4), 1>::smem_bytes();
            auto kern = bps_qr_kernel<L, M, R, P, 1, 4>;
            SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)A->ntiles, 32, smem, c->stream>>>(A->REC->p, F->GT->p, F->OUT->p, A->n, A->np, F->npo);
        } else if (look >= 2) {
            const size_t smem = WindowRing<D::NFA, bps_nst(L, M, 2), 1>::smem_bytes();
            auto kern = bps_qr_kernel<L, M, R, P, 1, 2>;
            SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)A->ntiles, 32, smem, c->stream>>>(A->REC->p, F->GT->p, F->OUT->p, A->n, A->np, F->npo);
        } else {
            const size_t smem = WindowRing<D::NFA, bps_nst(L, M, 1), 1>::smem_bytes();
            auto kern = bps_qr_kernel<L, M, R, P, 1, 1>;
            SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)A->ntiles, 32, smem, c->stream>>>(A->REC->p, F->GT->p, F->OUT->p, A->n, A->np, F->npo);
        }
    }
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}
template <int L, int M, int R, int P>
static int gram_shape(ssm_ctx *c, const ssm_bps *A) {
    using D = BpsDims<L, M, R, P>;
    bps_reduce_kernel<D::NFA, D::FU, R><<<(unsigned)A->ntiles, 32 * BPS_RW, 0, c->stream>>>(A->REC->p, nullptr, A->GT->p, nullptr, A->n, A->np, 0);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    A->gt_valid = true;
    return 0;
}
int launch_bps_gram(ssm_ctx *c, const ssm_bps *A) {
    if (!A->GT) SSM_TRY(dev_alloc(c, (size_t)(A->ntiles * A->tr * A->tr * 32) + 32, true, A->GT));
#define X(L_, M_, R_, P_) if (A->tl == L_ && A->tm == M_ && A->tr == R_ && A->tp == P_) return gram_shape<L_, M_, R_, P_>(c, A);
    SSM_BPS_SHAPES(X)
#undef X
    set_error("no BPS kernel for padded shape (%d,%d,%d,%d)", A->tl, A->tm, A->tr, A->tp);
    return SSM_E_UNSUPPORTED;
}
int launch_bps_qr(ssm_ctx *c, const ssm_bps *A, ssm_bpsqr *F) {
#define X(L_, M_, R_, P_) if (A->tl == L_ && A->tm == M_ && A->tr == R_ && A->tp == P_) return qr_shape<L_, M_, R_, P_>(c, A, F);
    SSM_BPS_SHAPES(X)
#undef X
    set_error("no BPS kernel for padded shape (%d,%d,%d,%d)", A->tl, A->tm, A->tr, A->tp);
    return SSM_E_UNSUPPORTED;
}

// Ring depth of the per-step solve kernels (batches too large for the grouped ring).  As for the AlmostBanded K2 / K3 (pick_stages, ab_kernels.cu):
// a batch that fits the SMs in one wave needs only what Little's law asks per warp (64 KB / warps per SM) — at config-3 shape and 65,536 systems
// (14 tiles per SM) two stages instead of six to eight take the solve from 10.7 to 9.7 ms; batches of several waves keep the deep ring (131,072
// systems: 19.4 ms deep, 22.0 ms shallow).
static int bps_stages(const ssm_ctx *c, size_t stage_bytes, int64_t ntiles) {
    if (c->pipeline_stages > 0) return c->pipeline_stages;
    const int64_t per_sm = (ntiles + c->sm_count - 1) / (c->sm_count > 0 ? c->sm_count : 148);
    size_t budget = 24000;
    if (per_sm >= 1 && per_sm <= 14 && (size_t)(64 * 1024) / (size_t)per_sm < budget) budget = (size_t)(64 * 1024) / (size_t)per_sm;
    int s = (int)((budget + stage_bytes / 2) / (stage_bytes + 8));
    return s < 2 ? 2 : (s > 8 ? 8 : s);
}

template <int L, int M, int R, int P>
static int qt_shape(ssm_ctx *c, const ssm_bpsqr *F, const double *rhs_in, double *rhs, int64_t npv) {
    using D = BpsDims<L, M, R, P>;
    bps_reduce_kernel<D::NFA, D::FU, R><<<(unsigned)F->ntiles, 32 * BPS_RW, 0, c->stream>>>(F->REC->p, rhs_in, nullptr, F->TT->p, F->n, F->np, npv);
    c->launches++;
    if ((c->ab_variant == 0 && F->ntiles <= (int64_t)c->sm_count * c->group_max_per_sm) || c->ab_variant == 4) {
        using RG = GroupRing<R, D::NFQ, 1, +1, 1, D::NFA, D::NFO, 1>;
        int ns, gm;
        if (pick_group(c, RG::BLOCK_BYTES, F->ntiles, &ns, &gm)) {
            const size_t smem = RG::smem_bytes(ns, gm);
            auto kern = bps_qt_group_kernel<L, M, R, P>;
            if (smem > 48 * 1024) SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            kern<<<(unsigned)F->ntiles, 64, smem, c->stream>>>(F->REC->p, F->OUT->p, F->GT->p, F->TT->p, rhs_in, rhs, F->n, F->np, F->npo, npv, ns, gm);
            c->launches++;
            SSM_CUDA(cudaGetLastError());
            return 0;
        }
    }
    if (c->ab_variant == 1) {
        bps_qt_kernel<L, M, R, P, DirectLoader><<<(unsigned)F->ntiles, 32, 0, c->stream>>>(F->REC->p, F->OUT->p, F->GT->p, F->TT->p, rhs_in, rhs, F->n,
                                                                                           F->np, F->npo, npv, 0);
    } else {
        using LD = RingLoader<R, D::NFQ, 1, +1, D::NFA, D::NFO, 1>;
        const int ns = bps_stages(c, LD::STAGE_BYTES, F->ntiles);
        const size_t smem = LD::smem_bytes(ns);
        auto kern = bps_qt_kernel<L, M, R, P, RingLoader>;
        if (smem > 48 * 1024) SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)F->ntiles, 32, smem, c->stream>>>(F->REC->p, F->OUT->p, F->GT->p, F->TT->p, rhs_in, rhs, F->n, F->np, F->npo, npv, ns);
    }
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}
template <int L, int M, int R, int P>
static int bs_shape(ssm_ctx *c, const ssm_bpsqr *F, const double *rhs_in, double *rhs, int64_t npv) {
    using D = BpsDims<L, M, R, P>;
    if ((c->ab_variant == 0 && F->ntiles <= (int64_t)c->sm_count * c->group_max_per_sm) || c->ab_variant == 4) {
        using RG = GroupRing<D::NFRP, P, 1, -1, 1, D::NFO, D::NFA, 1>;
        int ns, gm;
        if (pick_group(c, RG::BLOCK_BYTES, F->ntiles, &ns, &gm)) {
            const size_t smem = RG::smem_bytes(ns, gm);
            auto kern = bps_bs_group_kernel<L, M, R, P>;
            if (smem > 48 * 1024) SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            kern<<<(unsigned)F->ntiles, 64, smem, c->stream>>>(F->REC->p, F->OUT->p, rhs_in, rhs, F->n, F->np, F->npo, npv, ns, gm);
            c->launches++;
            SSM_CUDA(cudaGetLastError());
            return 0;
        }
    }
    if (c->ab_variant == 1) {
        bps_bs_kernel<L, M, R, P, DirectLoader><<<(unsigned)F->ntiles, 32, 0, c->stream>>>(F->REC->p, F->OUT->p, rhs_in, rhs, F->n, F->np, F->npo, npv, 0);
    } else {
        using LD = RingLoader<D::NFRP, P, 1, -1, D::NFO, D::NFA, 1>;
        const int ns = bps_stages(c, LD::STAGE_BYTES, F->ntiles);
        const size_t smem = LD::smem_bytes(ns);
        auto kern = bps_bs_kernel<L, M, R, P, RingLoader>;
        if (smem > 48 * 1024) SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)F->ntiles, 32, smem, c->stream>>>(F->REC->p, F->OUT->p, rhs_in, rhs, F->n, F->np, F->npo, npv, ns);
    }
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}
int launch_bps_qt(ssm_ctx *c, const ssm_bpsqr *F, const double *rhs_in, double *rhs) {
    const int64_t npv = (int64_t)F->n + PAD;
#define X(L_, M_, R_, P_) if (F->tl == L_ && F->tm == M_ && F->tr == R_ && F->tp == P_) return qt_shape<L_, M_, R_, P_>(c, F, rhs_in, rhs, npv);
    SSM_BPS_SHAPES(X)
#undef X
    return SSM_E_UNSUPPORTED;
}
int launch_bps_backsolve(ssm_ctx *c, const ssm_bpsqr *F, const double *rhs_in, double *rhs) {
    const int64_t npv = (int64_t)F->n + PAD;
#define X(L_, M_, R_, P_) if (F->tl == L_ && F->tm == M_ && F->tr == R_ && F->tp == P_) return bs_shape<L_, M_, R_, P_>(c, F, rhs_in, rhs, npv);
    SSM_BPS_SHAPES(X)
#undef X
    return SSM_E_UNSUPPORTED;
}
int launch_bps_mul(ssm_ctx *c, const ssm_bps *A, const double *x, double *y, const double *b, double *relres) {
    bps_mul_kernel<<<(unsigned)A->ntiles, 32, 0, c->stream>>>(A->REC->p, x, y, b, relres, shape_of(A), A->nb, A->np, (int64_t)A->n + PAD);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}
int launch_bps_upper_ldiv(ssm_ctx *c, const ssm_bps *A, double *rhs) {
    bps_upper_generic_kernel<<<(unsigned)A->ntiles, 32, 0, c->stream>>>(A->REC->p, rhs, shape_of(A), A->np, (int64_t)A->n + PAD);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ssm
