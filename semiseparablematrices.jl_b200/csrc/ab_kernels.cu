// ab_kernels.cu — sm_100a kernels of the AlmostBanded hot path (K1 ab_qr, K2 ab_qt, K3 ab_backsolve) and their
// dispatch.  One lane per system, one warp (= one CTA) per tile of 32 systems; the n axis is swept sequentially
// with the Householder window held in registers (ab_core.cuh) and the record streams pulled either by direct
// loads or by the bulk-async (TMA) shared-memory ring (stream_loader.cuh).
//
// Replaces: _almostbanded_qr! (AlmostBandedMatrix.jl:135-172) incl. the widening copy (:30-38),
//           lmul!(Q',b) (:181,189), _almostbanded_upper_ldiv! (:215-240).
#include <type_traits>
#include "ab_core.cuh"
#include "ab_generic.cuh"
#include "ssm_internal.cuh"
#include "stream_loader.cuh"

namespace ssm {

// ---------------------------------------------------------------------------------------------------------
// K1: QR sweep.  RHS: also carry b through the reflectors (fused Q'b, used by A\b).  STOREQ: keep reflectors.
// ---------------------------------------------------------------------------------------------------------
template <int L, int U, int R, bool RHS, bool STOREQ, template <int, int, int, int> class Loader>
__global__ void __launch_bounds__(32) ab_qr_kernel(const AbArgs a, const int nstage) {
    using D = AbDims<L, U, R>;
    constexpr int P = D::P, UB = D::UB, NFA = D::NFA, NFR = D::NFR, NFQ = D::NFQ;
    extern __shared__ __align__(128) unsigned char smem[];
    const int64_t tile = blockIdx.x;
    const int lane = threadIdx.x;
    const int n = a.n;
    const double *au = a.AU + tile * a.np * NFA * 32 + lane;
    const double *vv = a.V + tile * a.np * (R > 0 ? R : 1) * 32 + lane;
    double *rhs = RHS ? a.rhs + tile * a.np * 32 + lane : nullptr;
    const double *rin = RHS ? a.rhs_in + tile * a.np * 32 + lane : nullptr;
    double *ru = a.RU + tile * a.np * NFR * 32 + lane;
    double *qt = STOREQ ? a.QT + tile * a.np * NFQ * 32 + lane : nullptr;

    AbState<L, U, R> st;
    ab_prologue<L, U, R, RHS>(
        st, [&](int rho, int f) { return __ldg(au + (rho * NFA + f) * 32); },
        [&](int c, int q) { return __ldg(vv + (c * R + q) * 32); }, [&](int rho) { return rin[rho * 32]; });

    const int64_t nsteps = ((n + P - 1) / P) * P;
    Loader<NFA, R, RHS ? 1 : 0, +1> ld;
    ld.init(au + (int64_t)L * NFA * 32, vv + (int64_t)(UB + 1) * R * 32, RHS ? rin + L * 32 : nullptr, smem, nstage, nsteps);

    for (int kb = 0; kb < n; kb += P) {
#pragma unroll
        for (int ph = 0; ph < P; ++ph) {
            const int k = kb + ph;
            double rec[NFA], vc[R > 0 ? R : 1], bb[1];
            ld.acquire(k, rec, vc, bb);
            double out_r[NFR], out_q[NFQ], out_c;
            ab_step<L, U, R, RHS>(st, ph, rec, vc, bb[0], k < a.ncols, out_r, out_q, out_c);
            if (k < n) {
#pragma unroll
                for (int f = 0; f < NFR; ++f) ru[((int64_t)k * NFR + f) * 32] = out_r[f];
                if (STOREQ) {
#pragma unroll
                    for (int f = 0; f < NFQ; ++f) qt[((int64_t)k * NFQ + f) * 32] = out_q[f];
                }
                if (RHS) rhs[(int64_t)k * 32] = out_c;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K2: b <- Q'b from stored reflectors.
// ---------------------------------------------------------------------------------------------------------
template <int L, template <int, int, int, int> class Loader>
__global__ void __launch_bounds__(32) ab_qt_kernel(const AbSolveArgs a, const int nstage) {
    constexpr int P = L + 1, NFQ = L + 1;
    extern __shared__ __align__(128) unsigned char smem[];
    const int64_t tile = blockIdx.x;
    const int lane = threadIdx.x;
    const int n = a.n;
    const double *qt = a.QT + tile * a.np * NFQ * 32 + lane;
    double *rhs = a.rhs + tile * a.np * 32 + lane;
    const double *rin = a.rhs_in + tile * a.np * 32 + lane;
    QtState<L> st;
    qt_prologue<L>(st, [&](int rho) { return rin[rho * 32]; });
    const int64_t nsteps = ((n + P - 1) / P) * P;
    Loader<NFQ, 1, 0, +1> ld;
    ld.init(qt, rin + L * 32, nullptr, smem, nstage, nsteps);
    for (int kb = 0; kb < n; kb += P) {
#pragma unroll
        for (int ph = 0; ph < P; ++ph) {
            const int k = kb + ph;
            double q[NFQ], bb[1], dummy[1];
            ld.acquire(k, q, bb, dummy);
            const double c = qt_step<L>(st, ph, q, bb[0]);
            if (k < n) rhs[(int64_t)k * 32] = c;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K3: x <- R \ c, rows n-1 .. 0.
// ---------------------------------------------------------------------------------------------------------
template <int L, int U, int R, template <int, int, int, int> class Loader>
__global__ void __launch_bounds__(32) ab_backsolve_kernel(const AbSolveArgs a, const int nstage) {
    using D = AbDims<L, U, R>;
    constexpr int UB = D::UB, PX = UB + 1, NFR = D::NFR;
    extern __shared__ __align__(128) unsigned char smem[];
    const int64_t tile = blockIdx.x;
    const int lane = threadIdx.x;
    const int n = a.n;
    const double *ru = a.RU + tile * a.np * NFR * 32 + lane;
    const double *vv = a.V + tile * a.np * (R > 0 ? R : 1) * 32 + lane;
    double *rhs = a.rhs + tile * a.np * 32 + lane;
    const double *rin = a.rhs_in + tile * a.np * 32 + lane;
    BsState<UB, R> st;
    bs_init<UB, R>(st);
    const int nblk = (n + PX - 1) / PX;
    const int itop = nblk * PX - 1;                 // first (padded) row visited; rows >= n are skipped
    const int64_t nsteps = (int64_t)nblk * PX;
    Loader<NFR, R, 1, -1> ld;
    ld.init(ru + (int64_t)itop * NFR * 32, vv + (int64_t)(itop + UB + 1) * R * 32, rin + (int64_t)itop * 32, smem, nstage, nsteps);
    const bool unit = (a.unit & 1) != 0;
    int64_t g = 0;
    for (int ib = (nblk - 1) * PX; ib >= 0; ib -= PX) {
#pragma unroll
        for (int ph = PX - 1; ph >= 0; --ph) {
            const int i = ib + ph;
            double r[NFR], vc[R > 0 ? R : 1], cc[1];
            ld.acquire(g++, r, vc, cc);
            if (i < n) {
                const double x = bs_step<UB, R>(st, ph, r, vc, cc[0], unit);
                rhs[(int64_t)i * 32] = x;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Small-batch variants of K1/K2/K3 (GroupRing, stream_loader.cuh): one consumer warp per tile plus one producer warp that
// feeds the ring.  Same arithmetic (ab_core.cuh), same records, same results bit for bit; only the loader cost per step
// changes.  (Splitting a tile's 32 systems over 2 or 4 consumer warps was built and measured: no gain — a sweep is bound by
// its serial chain of n steps, which has the same length however few systems a warp carries.)
// ---------------------------------------------------------------------------------------------------------
template <int L, int U, int R, bool RHS, bool STOREQ>
__global__ void __launch_bounds__(64) ab_qr_group_kernel(const AbArgs a, const int nstage, const int gm) {
    using D = AbDims<L, U, R>;
    constexpr int P = D::P, UB = D::UB, NFA = D::NFA, NFR = D::NFR, NFQ = D::NFQ;
    using Ring = GroupRing<NFA, R, RHS ? 1 : 0, +1, P>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int64_t tile = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = a.n;
    const int64_t nblocks = (n + P - 1) / P;
    const double *au_t = a.AU + tile * a.np * NFA * 32;
    const double *vv_t = a.V + tile * a.np * (R > 0 ? R : 1) * 32;
    const double *rin_t = RHS ? a.rhs_in + tile * a.np * 32 : nullptr;
    Ring::init_barriers(smem, nstage, gm, 1);
    if (warp == 1) {
        if (lane == 0)
            Ring::produce(au_t + (int64_t)L * NFA * 32, vv_t + (int64_t)(UB + 1) * R * 32, RHS ? rin_t + L * 32 : nullptr, smem, nstage, gm, nblocks);
        return;
    }
    const int col = lane;
    const double *au = au_t + col, *vv = vv_t + col, *rin = RHS ? rin_t + col : nullptr;
    double *rhs = RHS ? a.rhs + tile * a.np * 32 + col : nullptr;
    double *ru = a.RU + tile * a.np * NFR * 32 + col;
    double *qt = STOREQ ? a.QT + tile * a.np * NFQ * 32 + col : nullptr;

    AbState<L, U, R> st;
    ab_prologue<L, U, R, RHS>(
        st, [&](int rho, int f) { return __ldg(au + (rho * NFA + f) * 32); },
        [&](int c, int q) { return __ldg(vv + (c * R + q) * 32); }, [&](int rho) { return rin[rho * 32]; });
    Ring ld;
    ld.attach(smem, nstage, gm, nblocks, col);
    // one block of P columns; only the last block of a sweep can reach past row n-1, so every other block runs without the
    // per-column bounds test (and the branch that comes with it)
    auto block = [&](const int kb, auto guarded) {
        constexpr bool GUARD = decltype(guarded)::value;
        ld.block_begin();
#pragma unroll
        for (int ph = 0; ph < P; ++ph) {
            const int k = kb + ph;
            double rec[NFA], vc[R > 0 ? R : 1], bb[1];
            ld.read(ph, rec, vc, bb);
            double out_r[NFR], out_q[NFQ], out_c;
            ab_step<L, U, R, RHS>(st, ph, rec, vc, bb[0], k < a.ncols, out_r, out_q, out_c);
            if (!GUARD || k < n) {
#pragma unroll
                for (int f = 0; f < NFR; ++f) ru[((int64_t)k * NFR + f) * 32] = out_r[f];
                if (STOREQ) {
#pragma unroll
                    for (int f = 0; f < NFQ; ++f) qt[((int64_t)k * NFQ + f) * 32] = out_q[f];
                }
                if (RHS) rhs[(int64_t)k * 32] = out_c;
            }
        }
        ld.block_end();
    };
    int kb = 0;
    for (; kb + P <= n; kb += P) block(kb, std::false_type{});
    if (kb < n) block(kb, std::true_type{});
}

template <int L>
__global__ void __launch_bounds__(64) ab_qt_group_kernel(const AbSolveArgs a, const int nstage, const int gm) {
    constexpr int P = L + 1, NFQ = L + 1;
    using Ring = GroupRing<NFQ, 1, 0, +1, P>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int64_t tile = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = a.n;
    const int64_t nblocks = (n + P - 1) / P;
    const double *qt_t = a.QT + tile * a.np * NFQ * 32;
    const double *rin_t = a.rhs_in + tile * a.np * 32;
    Ring::init_barriers(smem, nstage, gm, 1);
    if (warp == 1) {
        if (lane == 0) Ring::produce(qt_t, rin_t + L * 32, nullptr, smem, nstage, gm, nblocks);
        return;
    }
    const int col = lane;
    double *rhs = a.rhs + tile * a.np * 32 + col;
    const double *rin = rin_t + col;
    QtState<L> st;
    qt_prologue<L>(st, [&](int rho) { return rin[rho * 32]; });
    Ring ld;
    ld.attach(smem, nstage, gm, nblocks, col);
    auto block = [&](const int kb, auto guarded) {
        constexpr bool GUARD = decltype(guarded)::value;
        ld.block_begin();
#pragma unroll
        for (int ph = 0; ph < P; ++ph) {
            const int k = kb + ph;
            double q[NFQ], bb[1], dummy[1];
            ld.read(ph, q, bb, dummy);
            const double c = qt_step<L>(st, ph, q, bb[0]);
            if (!GUARD || k < n) rhs[(int64_t)k * 32] = c;
        }
        ld.block_end();
    };
    int kb = 0;
    for (; kb + P <= n; kb += P) block(kb, std::false_type{});
    if (kb < n) block(kb, std::true_type{});
}

template <int L, int U, int R>
__global__ void __launch_bounds__(64) ab_backsolve_group_kernel(const AbSolveArgs a, const int nstage, const int gm) {
    using D = AbDims<L, U, R>;
    constexpr int UB = D::UB, PX = UB + 1, NFR = D::NFR;
    using Ring = GroupRing<NFR, R, 1, -1, PX>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int64_t tile = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = a.n;
    const int nblk = (n + PX - 1) / PX;
    const int itop = nblk * PX - 1;
    const double *ru_t = a.RU + tile * a.np * NFR * 32;
    const double *vv_t = a.V + tile * a.np * (R > 0 ? R : 1) * 32;
    const double *rin_t = a.rhs_in + tile * a.np * 32;
    Ring::init_barriers(smem, nstage, gm, 1);
    if (warp == 1) {
        if (lane == 0)
            Ring::produce(ru_t + (int64_t)itop * NFR * 32, vv_t + (int64_t)(itop + UB + 1) * R * 32, rin_t + (int64_t)itop * 32, smem, nstage, gm, nblk);
        return;
    }
    const int col = lane;
    double *rhs = a.rhs + tile * a.np * 32 + col;
    BsState<UB, R> st;
    bs_init<UB, R>(st);
    Ring ld;
    ld.attach(smem, nstage, gm, nblk, col);
    const bool unit = a.unit != 0;
    // only the first (topmost) block can contain rows >= n
    auto block = [&](const int ib, auto guarded) {
        constexpr bool GUARD = decltype(guarded)::value;
        ld.block_begin();
#pragma unroll
        for (int ph = PX - 1; ph >= 0; --ph) {
            const int i = ib + ph;
            double r[NFR], vc[R > 0 ? R : 1], cc[1];
            ld.read(PX - 1 - ph, r, vc, cc);
            if (!GUARD || i < n) {
                const double x = bs_step<UB, R>(st, ph, r, vc, cc[0], unit);
                rhs[(int64_t)i * 32] = x;
            }
        }
        ld.block_end();
    };
    int ib = (nblk - 1) * PX;
    block(ib, std::true_type{});
    for (ib -= PX; ib >= 0; ib -= PX) block(ib, std::false_type{});
}

// ---------------------------------------------------------------------------------------------------------
// generic (runtime l,u,r) kernels — any shape within SSM_MAX_*; also Q*b and the operand-level triangular solves
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) ab_qr_generic_kernel(const AbArgs a, int l, int u, int r, int with_rhs, int store_q) {
    generic_ab_qr(a, l, u, r, with_rhs != 0, store_q != 0, (int64_t)blockIdx.x, (int)threadIdx.x);
}
__global__ void __launch_bounds__(32) ab_qt_generic_kernel(const AbSolveArgs a, int l, int transpose) {
    generic_ab_qt(a, l, transpose != 0, (int64_t)blockIdx.x, (int)threadIdx.x);
}
__global__ void __launch_bounds__(32) ab_backsolve_generic_kernel(const AbSolveArgs a, int l, int u, int r) {
    generic_ab_backsolve(a, l, u, r, (int64_t)blockIdx.x, (int)threadIdx.x);
}
__global__ void __launch_bounds__(32) ab_tri_generic_kernel(const double *AU, const double *V, double *rhs, int n, int64_t np,
                                                            int l, int u, int r, int upper, int unit) {
    generic_ab_tri(AU, V, rhs, n, np, l, u, r, upper != 0, unit != 0, (int64_t)blockIdx.x, (int)threadIdx.x);
}

// ---------------------------------------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------------------------------------
static int pick_stages(const ssm_ctx *c, size_t stage_bytes, int64_t ntiles) {
    if (c->pipeline_stages > 0) return c->pipeline_stages;
    // The ring is the only thing that hides HBM latency for a single-warp CTA, so it takes the shared memory the grid leaves
    // free: ~15 KB per warp when 14 CTAs per SM are co-resident (all 2048 tiles of the headline config), proportionally
    // deeper when the batch is small (Little: ~6.5 MB must be in flight device-wide to reach the HBM rate).
    int64_t per_sm = (ntiles + c->sm_count - 1) / (c->sm_count > 0 ? c->sm_count : 148);
    if (per_sm < 1) per_sm = 1;
    const bool waves = per_sm > 14;          // more tiles than an SM holds at once: the deep ring also caps residency at ~14 CTAs, which measured best there
    if (per_sm > 14) per_sm = 14;
    size_t budget = (size_t)(210 * 1024) / (size_t)per_sm;
    // ... but no deeper than the HBM rate needs: ~6.5 MB in flight device-wide is 44 KB / per_sm for each warp; 64 KB / per_sm leaves a margin.
    // Measured at config 2 (14 warps per SM): K3 with 2 stages instead of 5 and K2 with 4 instead of 14: ldiv! 1.349 -> 1.272 ms
    // (tools/scan_ldiv_stages.sh) — a ring that takes all the shared memory there is only costs L1 and re-arming work.  With more than 14 tiles
    // per SM the shallow ring lets ~28 CTAs in at once and ldiv! gets slower (131,072 systems: 2.51 -> 2.93 ms), so the deep ring stays there.
    const size_t want = (size_t)(64 * 1024) / (size_t)per_sm;
    if (budget > want && !waves) budget = want;
    int s = (int)((budget + (waves ? 0 : stage_bytes / 2)) / (stage_bytes + 8));
    if (s < 2) s = 2;
    if (s > 48) s = 48;
    return s;
}

// shared-memory attributes of a ring kernel before its launch
template <class K>
static int launch1(ssm_ctx *, K kern, int64_t, size_t smem, const char *) {
    if (smem > 48 * 1024) SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (smem > 0) SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return 0;
}

// GroupRing geometry for a grid of ntiles CTAs: as many stages of gm blocks as the shared memory left per CTA holds; returns
// false when not even two single-block stages fit (the caller then keeps the per-step ring).
bool pick_group(const ssm_ctx *c, size_t block_bytes, int64_t ntiles, int *nstage, int *gm) {
    int64_t per_sm = (ntiles + c->sm_count - 1) / (c->sm_count > 0 ? c->sm_count : 148);
    if (per_sm < 1) per_sm = 1;
    size_t budget = (size_t)(200 * 1024) / (size_t)per_sm;
    if (budget > (size_t)c->group_smem_kb * 1024) budget = (size_t)c->group_smem_kb * 1024;
    int g = (int)((size_t)c->group_bytes / block_bytes);   // aim at group_bytes per bulk copy group ...
    if (g < 1) g = 1;
    int s = (int)(budget / (g * block_bytes + 16));
    while (s < 6 && g > 1) { --g; s = (int)(budget / (g * block_bytes + 16)); }   // ... but keep at least 6 stages in flight
    if (s < 2) return false;
    if (s > 12) s = 12;
    if (c->pipeline_stages > 0) s = c->pipeline_stages;
    if ((size_t)s * (g * block_bytes + 16) > (size_t)220 * 1024) return false;
    *nstage = s; *gm = g;
    return true;
}

template <class K>
static int launch_group(ssm_ctx *c, K kern, size_t smem) {
    if (smem > 48 * 1024) SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    (void)c;
    return 0;
}

static inline bool is_group_variant(int v) { return v == 4; }

#define SSM_GROUP_GO(KERN_, ARGS_)                                                                 \
    {                                                                                              \
        auto kern = KERN_;                                                                         \
        SSM_TRY(launch_group(c, kern, smem));                                                      \
        kern<<<(unsigned)ntiles, 64, smem, c->stream>>>(ARGS_, ns, gm);                            \
    }

template <int L, int U, int R, bool RHS, bool SQ>
static int qr_group(ssm_ctx *c, int64_t ntiles, const AbArgs &a, bool *done) {
    using D = AbDims<L, U, R>;
    using RG = GroupRing<D::NFA, R, RHS ? 1 : 0, +1, D::P>;
    int ns, gm;
    *done = false;
    if (!pick_group(c, RG::BLOCK_BYTES, ntiles, &ns, &gm)) return 0;
    const size_t smem = RG::smem_bytes(ns, gm);
    SSM_GROUP_GO((ab_qr_group_kernel<L, U, R, RHS, SQ>), a)
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    *done = true;
    return 0;
}

template <int L, int U, int R>
static int qr_inst(ssm_ctx *c, int64_t ntiles, const AbArgs &a, bool with_rhs, bool store_q, int variant) {
    using D = AbDims<L, U, R>;
    dim3 grid((unsigned)ntiles), block(32);
    if (is_group_variant(variant)) {       // instantiated for qr (keep reflectors) and for the fused A\\b sweep
        bool done = false;
        if (!with_rhs && store_q) SSM_TRY((qr_group<L, U, R, false, true>(c, ntiles, a, &done)));
        if (with_rhs && !store_q) SSM_TRY((qr_group<L, U, R, true, false>(c, ntiles, a, &done)));
        if (done) return 0;
        variant = 2;
    }
#define SSM_QR_CASE(RHS_, SQ_)                                                                                   \
    if (with_rhs == RHS_ && store_q == SQ_) {                                                                    \
        if (variant == 2) {                                                                                      \
            using LD = RingLoader<D::NFA, R, RHS_ ? 1 : 0, +1>;                                                  \
            const int ns = pick_stages(c, LD::STAGE_BYTES, ntiles);                                                      \
            const size_t smem = LD::smem_bytes(ns);                                                              \
            auto kern = ab_qr_kernel<L, U, R, RHS_, SQ_, RingLoaderD>;                                            \
            SSM_TRY(launch1(c, kern, ntiles, smem, "ab_qr_ring"));                                               \
            kern<<<grid, block, smem, c->stream>>>(a, ns);                                                       \
        } else {                                                                                                 \
            auto kern = ab_qr_kernel<L, U, R, RHS_, SQ_, DirectLoaderD>;                                         \
            if (c->carveout_all) SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
            kern<<<grid, block, 0, c->stream>>>(a, 0);                                                           \
        }                                                                                                        \
    }
    SSM_QR_CASE(false, true)
    SSM_QR_CASE(true, true)
    SSM_QR_CASE(true, false)
    SSM_QR_CASE(false, false)
#undef SSM_QR_CASE
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}

template <int L>
static int qt_inst(ssm_ctx *c, int64_t ntiles, const AbSolveArgs &a, int variant) {
    dim3 grid((unsigned)ntiles), block(32);
    if (is_group_variant(variant)) {
        using RG = GroupRing<L + 1, 1, 0, +1, L + 1>;
        int ns, gm;
        if (pick_group(c, RG::BLOCK_BYTES, ntiles, &ns, &gm)) {
            const size_t smem = RG::smem_bytes(ns, gm);
            SSM_GROUP_GO((ab_qt_group_kernel<L>), a)
            c->launches++;
            SSM_CUDA(cudaGetLastError());
            return 0;
        }
        variant = 2;
    }
    if (variant == 2) {
        using LD = RingLoader<L + 1, 1, 0, +1>;
        const int ns = c->qt_stages > 0 ? c->qt_stages : pick_stages(c, LD::STAGE_BYTES, ntiles);
        const size_t smem = LD::smem_bytes(ns);
        auto kern = ab_qt_kernel<L, RingLoaderD>;
        SSM_TRY(launch1(c, kern, ntiles, smem, "ab_qt_ring"));
        kern<<<grid, block, smem, c->stream>>>(a, ns);
    } else {
        auto kern = ab_qt_kernel<L, DirectLoaderD>;
        if (c->carveout_all) SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        kern<<<grid, block, 0, c->stream>>>(a, 0);
    }
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}

template <int L, int U, int R>
static int bs_inst(ssm_ctx *c, int64_t ntiles, const AbSolveArgs &a, int variant) {
    using D = AbDims<L, U, R>;
    dim3 grid((unsigned)ntiles), block(32);
    if (is_group_variant(variant)) {
        using RG = GroupRing<D::NFR, R, 1, -1, D::UB + 1>;
        int ns, gm;
        if (pick_group(c, RG::BLOCK_BYTES, ntiles, &ns, &gm)) {
            const size_t smem = RG::smem_bytes(ns, gm);
            SSM_GROUP_GO((ab_backsolve_group_kernel<L, U, R>), a)
            c->launches++;
            SSM_CUDA(cudaGetLastError());
            return 0;
        }
        variant = 2;
    }
    if (variant == 2) {
        using LD = RingLoader<D::NFR, R, 1, -1>;
        const int ns = c->bs_stages > 0 ? c->bs_stages : pick_stages(c, LD::STAGE_BYTES, ntiles);
        const size_t smem = LD::smem_bytes(ns);
        auto kern = ab_backsolve_kernel<L, U, R, RingLoaderD>;
        SSM_TRY(launch1(c, kern, ntiles, smem, "ab_bs_ring"));
        kern<<<grid, block, smem, c->stream>>>(a, ns);
    } else {
        auto kern = ab_backsolve_kernel<L, U, R, DirectLoaderD>;
        if (c->carveout_all) SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        kern<<<grid, block, 0, c->stream>>>(a, 0);
    }
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}

// shapes with a register-resident fast path (README (1,0,1); tests (1,1,1),(2,0,2),(1,0,2); C2 (2,2,2); C4 (4,4,2))
#define SSM_AB_SHAPES(X) X(1, 0, 1) X(1, 1, 1) X(1, 0, 2) X(2, 0, 2) X(2, 1, 2) X(2, 2, 1) X(2, 2, 2) X(3, 3, 2) X(3, 3, 3) X(4, 4, 2)

// auto (0): measured on B200 at config 2 — the QR sweep is fastest with direct loads (it is register-rich and
// store-heavy), the two light solve kernels with the bulk-async ring (profiles/README.md)
// Below ~7 single-warp CTAs per SM there are too few warps to hide HBM latency behind one-step-ahead register loads, and the ring's
// deep bulk-async prefetch wins for the QR sweep as well (measured: n=1024 x 16,384 systems 0.81 -> 0.47 ms, n=8192 x 4,096 5.1 -> 3.6 ms).
// `earlier`: tiles per SM by which a kernel hands over to its big-batch form before the common threshold.  K2 / K3 do so two tiles per SM before K1
// since their per-step rings follow Little's law (40,960 systems of the config-2 shape: ldiv! 0.810 ms against 0.849 ms with the grouped ring).
static int effective_variant(const ssm_ctx *c, int kernel_default, int64_t ntiles, int earlier = 0) {
    if (c->ab_variant != 0) return c->ab_variant;
    if (ntiles <= (int64_t)c->sm_count * (c->group_max_per_sm - earlier)) return 4;   // small batch: GroupRing kernels
    return kernel_default;
}

int launch_ab_qr(ssm_ctx *c, int l, int u, int r, int64_t ntiles, const AbArgs &a, bool with_rhs, bool store_q) {
    if (ntiles <= 0) return 0;
    const int variant = effective_variant(c, 1, ntiles);
    if (variant != 9) {
#define X(L_, U_, R_) \
    if (l == L_ && u == U_ && r == R_) return qr_inst<L_, U_, R_>(c, ntiles, a, with_rhs, store_q, variant);
        SSM_AB_SHAPES(X)
#undef X
    }
    if (l > SSM_MAX_BW || u > SSM_MAX_BW || r > SSM_MAX_RANK) { set_error("(l,u,r)=(%d,%d,%d) exceeds compiled limits", l, u, r); return SSM_E_UNSUPPORTED; }
    ab_qr_generic_kernel<<<(unsigned)ntiles, 32, 0, c->stream>>>(a, l, u, r, with_rhs, store_q);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}

__global__ void __launch_bounds__(32) ab_qr_continue_kernel(const AbArgs a, int l, int u, int r, int kstart) {
    generic_ab_qr(a, l, u, r, false, true, (int64_t)blockIdx.x, (int)threadIdx.x, kstart);
}
// Register-resident second entry of K1 for the compiled shapes: the sweep re-enters at column kstart.  Everything the sweep addresses is
// relative to its first row, so the continuation is the SAME sweep on pointers moved down by kstart rows (row kstart plays row 0: it lives in
// slot 0, the phases line up), with a prologue that takes the l partially updated rows out of the factor records — RU: the diagonal and what is
// right of it plus the U row; QT of the columns kstart .. kstart+l-2: the entries below their diagonals — instead of out of the operand.
template <int L, int U, int R>
__global__ void __launch_bounds__(32) ab_qr_cont_kernel(const AbArgs a, const int kstart) {
    using D = AbDims<L, U, R>;
    constexpr int P = D::P, UB = D::UB, NFA = D::NFA, NFR = D::NFR, NFQ = D::NFQ;
    const int64_t tile = blockIdx.x;
    const int lane = threadIdx.x;
    const int n = a.n - kstart, ncols = a.ncols - kstart;
    const double *au = a.AU + (tile * a.np + kstart) * NFA * 32 + lane;
    const double *vv = a.V + (tile * a.np + kstart) * (R > 0 ? R : 1) * 32 + lane;
    double *ru = a.RU + (tile * a.np + kstart) * NFR * 32 + lane;
    double *qt = a.QT + (tile * a.np + kstart) * NFQ * 32 + lane;
    AbState<L, U, R> st;
#pragma unroll
    for (int i = 0; i < L; ++i) {
#pragma unroll
        for (int dd = 0; dd < D::ND; ++dd) st.A[i][dd] = 0.0;
#pragma unroll
        for (int t = 0; t <= UB; ++t)                 // entry (row i, column t) sits at diagonal index t - i + L
            st.A[i][t - i + L] = (t >= i) ? ru[(i * NFR + (t - i)) * 32] : qt[(t * NFQ + (i - t - 1)) * 32];
#pragma unroll
        for (int q = 0; q < R; ++q) st.UC[i][q] = ru[(i * NFR + UB + 1 + q) * 32];
        st.RH[i] = 0.0;
    }
    const int64_t nsteps = ((n + P - 1) / P) * P;
    DirectLoaderD<NFA, R, 0, +1> ld;
    ld.init(au + (int64_t)L * NFA * 32, vv + (int64_t)(UB + 1) * R * 32, nullptr, nullptr, 0, nsteps);
    for (int kb = 0; kb < n; kb += P) {
#pragma unroll
        for (int ph = 0; ph < P; ++ph) {
            const int k = kb + ph;
            double rec[NFA], vc[R > 0 ? R : 1], bb[1];
            ld.acquire(k, rec, vc, bb);
            double out_r[NFR], out_q[NFQ], out_c;
            ab_step<L, U, R, false>(st, ph, rec, vc, 0.0, k < ncols, out_r, out_q, out_c);
            if (k < n) {
#pragma unroll
                for (int f = 0; f < NFR; ++f) ru[((int64_t)k * NFR + f) * 32] = out_r[f];
#pragma unroll
                for (int f = 0; f < NFQ; ++f) qt[((int64_t)k * NFQ + f) * 32] = out_q[f];
            }
        }
    }
}
// continue a partial factorisation in place from column kstart (a.RU / a.QT hold it) to a.ncols reflectors
int launch_ab_qr_continue(ssm_ctx *c, int l, int u, int r, int64_t ntiles, const AbArgs &a, int kstart) {
    if (ntiles <= 0) return 0;
    if (c->ab_variant != 9) {
#define X(L_, U_, R_)                                                                              \
    if (l == L_ && u == U_ && r == R_) {                                                           \
        ab_qr_cont_kernel<L_, U_, R_><<<(unsigned)ntiles, 32, 0, c->stream>>>(a, kstart);          \
        c->launches++;                                                                             \
        SSM_CUDA(cudaGetLastError());                                                              \
        return 0;                                                                                  \
    }
        SSM_AB_SHAPES(X)
#undef X
    }
    ab_qr_continue_kernel<<<(unsigned)ntiles, 32, 0, c->stream>>>(a, l, u, r, kstart);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}

__global__ void __launch_bounds__(32) ab_grow_fix_kernel(const AbArgs a, int l, int u, int r, int n0, int k0) {
    generic_ab_grow_fix(a, l, u, r, n0, k0, (int64_t)blockIdx.x, (int)threadIdx.x);
}
// ssm_abqr_resize: re-derive the factor entries that reach into the columns >= n0 of the grown operand (ab_generic.cuh).  At most 2l+u steps
// of a runtime-shape replay per system: an edge computation, the same kernel for every shape.
int launch_ab_grow_fix(ssm_ctx *c, int l, int u, int r, int64_t ntiles, const AbArgs &a, int n0, int k0) {
    if (ntiles <= 0) return 0;
    ab_grow_fix_kernel<<<(unsigned)ntiles, 32, 0, c->stream>>>(a, l, u, r, n0, k0);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}

int launch_ab_qt(ssm_ctx *c, int l, int u, int r, int64_t ntiles, const AbSolveArgs &a, bool transpose) {
    (void)u; (void)r;
    if (ntiles <= 0) return 0;
    const int variant = effective_variant(c, 2, ntiles, 2);
    if (variant != 9 && transpose) {
        switch (l) {
            case 1: return qt_inst<1>(c, ntiles, a, variant);
            case 2: return qt_inst<2>(c, ntiles, a, variant);
            case 3: return qt_inst<3>(c, ntiles, a, variant);
            case 4: return qt_inst<4>(c, ntiles, a, variant);
            default: break;
        }
    }
    if (l > SSM_MAX_BW) { set_error("l=%d exceeds compiled limits", l); return SSM_E_UNSUPPORTED; }
    ab_qt_generic_kernel<<<(unsigned)ntiles, 32, 0, c->stream>>>(a, l, transpose);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}

int launch_ab_backsolve(ssm_ctx *c, int l, int u, int r, int64_t ntiles, const AbSolveArgs &a) {
    if (ntiles <= 0) return 0;
    const int variant = effective_variant(c, 2, ntiles, 2);
    if (variant != 9) {
#define X(L_, U_, R_) \
    if (l == L_ && u == U_ && r == R_) return bs_inst<L_, U_, R_>(c, ntiles, a, variant);
        SSM_AB_SHAPES(X)
#undef X
    }
    if (l > SSM_MAX_BW || u > SSM_MAX_BW || r > SSM_MAX_RANK) { set_error("(l,u,r)=(%d,%d,%d) exceeds compiled limits", l, u, r); return SSM_E_UNSUPPORTED; }
    ab_backsolve_generic_kernel<<<(unsigned)ntiles, 32, 0, c->stream>>>(a, l, u, r);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}

int launch_ab_tri(ssm_ctx *c, const ssm_ab *A, double *rhs, bool upper, bool unit) {
    if (A->ntiles <= 0) return 0;
    ab_tri_generic_kernel<<<(unsigned)A->ntiles, 32, 0, c->stream>>>(A->AU->p, A->V->p, rhs, A->n, A->np, A->l, A->u, A->r, upper, unit);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ssm

