// ssm_internal.cuh — shared definitions of libssmb200.so (not installed; the public surface is include/ssm_b200.h)
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <memory>
#include <string>
#include <vector>
#include "../../include/ssm_b200.h"

struct ssm_ctx;

namespace ssm {

// ---- device layout ------------------------------------------------------------------
// Systems are grouped in tiles of 32 (one warp, one lane per system).  Every per-row quantity is a
// "record": NF fields x 32 lanes of fp64, lane fastest, so one field of one record is one 256-byte
// coalesced segment and one record is NF*256 contiguous bytes.  A tile's records are contiguous:
//     addr(tile, rec, field, lane) = ((tile * NP + rec) * NF + field) * 32 + lane
// NP = n + PAD records per tile; the PAD tail is zero so look-ahead loads and the rows "entering" after
// the last row need no bounds checks.
constexpr int TILE = 32;
constexpr int PAD  = 48;

inline int64_t ntiles_of(int64_t nb) { return (nb + TILE - 1) / TILE; }

// The stream a context enqueues on, with a flag that outlives the context: a block may only be asked "is your stream idle?" while that
// stream still exists (cudaStreamQuery on a destroyed stream is undefined).  ssm_ctx_destroy synchronises, then clears `alive`.
struct StreamToken {
    cudaStream_t st = nullptr;
    std::atomic<bool> alive{true};
};
using StreamTokenPtr = std::shared_ptr<StreamToken>;

struct DevBuf {
    double *p = nullptr;
    size_t  count = 0;          // doubles asked for
    size_t  cap = 0;            // doubles the block holds (>= count when it came off the free list, api.cu)
    int     device = 0;
    StreamTokenPtr tok;         // stream of the context it was allocated for
    std::atomic<bool> foreign{false};   // some OTHER context's stream has read or written it (see uses()): never recycled, cudaFree waits for the device
    ~DevBuf();
};
using DevBufPtr = std::shared_ptr<DevBuf>;
int dev_alloc(ssm_ctx *c, size_t count, bool zero, DevBufPtr &out);

void set_error(const char *fmt, ...);
bool is_device_ptr(const void *p);
int  cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define SSM_CUDA(call)                                                      \
    do {                                                                    \
        cudaError_t e__ = (call);                                           \
        if (e__ != cudaSuccess) return ::ssm::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)
#define SSM_TRY(call)                        \
    do {                                     \
        int rc__ = (call);                   \
        if (rc__ != 0) return rc__;          \
    } while (0)

struct Staging {        // lazily grown scratch: device staging in reference layout + pinned host bounce
    double *dev = nullptr;  size_t dev_count = 0;
    double *pin = nullptr;  size_t pin_count = 0;
};

}  // namespace ssm

namespace ssm { int staging_dev(ssm_ctx *c, size_t count, double **out); void hostpipe_free(ssm_ctx *c); void abc_cache_free(ssm_ctx *c); }

struct ssm_ctx {
    int          device = 0;
    cudaStream_t stream = nullptr;
    bool         own_stream = false;
    ssm::StreamTokenPtr tok;           // {stream, alive}: what this context's device blocks ask before they go back to the free list
    std::atomic<int64_t> live{0};      // objects created on this context and not yet destroyed (ssm::CtxRef)
    std::atomic<bool>    retired{false};   // ssm_ctx_destroy was called while objects were alive: freed with the last of them
    int          ab_variant = 0;       // 0 auto, 1 direct-load kernels, 2 bulk-async (TMA) ring kernels, 4 GroupRing kernels (producer warp + grouped bulk copies), 9 generic
    int          carveout_all = 0;     // experiment: direct-load kernels also ask for the max-shared carveout
    int          group_smem_kb = 200;  // GroupRing: shared memory one CTA may take.  Alone on the GPU the deep ring is worth 3-8 % at <= 2 tiles per
                                       // SM; contexts that share the GPU on different streams set ~96 so that their CTAs fit side by side
    int          group_bytes = 32768;  // GroupRing: target size of one bulk-copy group
    int          group_max_per_sm = 9; // auto: batches of at most this many tiles per SM run the GroupRing kernels
    int          pipeline_stages = 0;  // 0 = kernel default
    int          qt_stages = 0, bs_stages = 0;   // ring depth of K2 / K3 alone (0 = pick_stages)
    int          abc_tiled_min_chunks = 12288;  // chunked path (K10): from this many chunks on stage 3 runs one lane per chunk over tiled records
    int          abc_lpc = 1;                   // chunked path, (l+u, r) = (8, 2) at many chunks: one lane per chunk in stage 1 (abc_qr_lpc_kernel): 0 never, 1 from the second solve of the same operands on, 2 always
    int          abc_gs = 4;                    // chunked path: lanes that share one chunk in stage 1 of the wide shapes (4, or 8 as for all others)
    int          sm_count = 148;
    int64_t      launches = 0;
    cudaEvent_t  ev = nullptr;         // ordering against other contexts' streams (ssm::CrossOrder), created on first use
    ssm::Staging stg;
    // extra streams / events for the chunk-pipelined host path
    cudaStream_t aux[3] = {nullptr, nullptr, nullptr};
    void        *hostpipe = nullptr;   // ssm::HostPipe of ssm_ab_solve_host (host_path.cu)
    void        *abc_cache[2] = {nullptr, nullptr};   // recently destroyed ssm_abc objects kept for reuse (abc_kernels.cu)
};

namespace ssm {
// Every object counts against its context.  A context destroyed before its objects (a host runtime that runs finalizers in any order, e.g. Julia's
// atexit hooks before its GC finalizers) is only retired — streams synchronised and released, caches freed — and its memory goes with the last object.
void ctx_retain(ssm_ctx *c);
void ctx_release(ssm_ctx *c);
struct CtxRef {
    ssm_ctx *c = nullptr;
    CtxRef() = default;
    CtxRef(const CtxRef &) = delete;
    CtxRef &operator=(const CtxRef &) = delete;
    void bind(ssm_ctx *ctx) { if (!c) { c = ctx; ctx_retain(ctx); } }
    ~CtxRef() { if (c) ctx_release(c); }
};
// wait for the object's context before its buffers go away (nothing to wait for once the context is retired: it synchronised then)
inline void obj_sync(const ssm_ctx *c) { if (c && !c->retired.load()) cudaStreamSynchronize(c->stream); }
// `user` is about to enqueue work that touches block d.  If d belongs to another context, the owner's destroy-time synchronisation no longer
// covers every use, so the block is barred from the free list for good (its destructor falls back to cudaFree, which waits for the device).
inline void uses(const DevBufPtr &d, const ssm_ctx *user) { if (d && d->tok.get() != user->tok.get()) d->foreign.store(true); }
// An entry point of context `user` is about to enqueue work that reads or writes objects of OTHER contexts (a vector made on context 2 solved with
// a factorisation of context 1).  Constructor: user's stream waits for everything the owners have enqueued so far (the operand may still be being
// written); destructor: every owner's stream waits for what user enqueued in between, so the owner's later calls (ssm_vec_get ...) see the result.
// Device-side ordering only (events), no host synchronisation.
struct CrossOrder {
    ssm_ctx *user; ssm_ctx *owners[4]; int n = 0;
    static void wait(ssm_ctx *waiter, ssm_ctx *signaller) {
        if (!signaller->ev && cudaEventCreateWithFlags(&signaller->ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return; }
        cudaEventRecord(signaller->ev, signaller->stream);
        cudaStreamWaitEvent(waiter->stream, signaller->ev, 0);
    }
    CrossOrder(ssm_ctx *u, std::initializer_list<ssm_ctx *> os) : user(u) {
        for (ssm_ctx *o : os) {
            if (!o || o == u || o->stream == u->stream || o->retired.load() || n == 4) continue;
            bool dup = false;
            for (int i = 0; i < n; ++i) dup = dup || owners[i] == o;
            if (dup) continue;
            owners[n++] = o;
            wait(u, o);
        }
    }
    ~CrossOrder() { for (int i = 0; i < n; ++i) wait(owners[i], user); }
    CrossOrder(const CrossOrder &) = delete;
    CrossOrder &operator=(const CrossOrder &) = delete;
};
}  // namespace ssm

struct ssm_vec {
    ssm_ctx *ctx; ssm::CtxRef ref; int n; int64_t nb, ntiles, np;
    ssm::DevBufPtr d;           // [ntiles][np][32]
};

struct ssm_ab {
    ssm_ctx *ctx; ssm::CtxRef ref; int n, l, u, r; int64_t nb, ntiles, np;
    ssm::DevBufPtr AU;          // [ntiles][np][l+u+1+r][32]   row record: A[i, i-l .. i+u], U[i, :]
    ssm::DevBufPtr V;           // [ntiles][np][r][32]         V[:, c]
    mutable ssm::DevBufPtr scratch;  // R records of the fused A\\b path (kept between calls)
};

struct ssm_abqr {
    ssm_ctx *ctx; ssm::CtxRef ref; int n, l, u, r; int64_t nb, ntiles, np;
    int ncols = 0;              // reflectors formed (n-1 for qr(A); fewer after ssm_ab_qr_cols)
    ssm::DevBufPtr RU;          // [ntiles][np][l+u+1+r][32]   R[k, k .. k+ub], Ufinal[k, :]
    ssm::DevBufPtr QT;          // [ntiles][np][l+1][32]       reflector tail of column k, tau_k
    ssm::DevBufPtr V;           // shared with the operand batch (never written by the path)
};

// BandedPlusSemiseparable operands.  (l,m,r,p) is the caller's shape; (tl,tm,tr,tp) >= it is the compiled kernel
// shape the batch is zero-padded to (a BPS matrix with zero extra bands / generator columns is the same matrix).
struct ssm_bps {
    ssm_ctx *ctx; ssm::CtxRef ref; int n, l, m, r, p, tl, tm, tr, tp; int64_t nb, ntiles, np;
    ssm::DevBufPtr REC;         // [ntiles][np][nfa][32], record (FRONT + row): B0 row (tl+tm+1) | U (tr) | V (tr) | W (tp) | S (tp)
    // U'U of every system ([ntiles][tr*tr][32]), the one reduction the QR sweep needs before its first step (semiseparableqr.jl:36-48,573-640).
    // It depends on the operands only, so it is formed when they are written (ssm_bps_set / _generate) and kept with them: qr(A) then makes
    // ONE pass over the records.  gt_valid is false for a batch written by another kernel (rq_mul's result); qr forms it on demand then.
    mutable ssm::DevBufPtr GT;
    mutable bool gt_valid = false;
};
struct ssm_bpsqr {
    ssm_ctx *ctx; ssm::CtxRef ref; int n, l, m, r, p, tl, tm, tr, tp; int64_t nb, ntiles, np, npo;
    ssm::DevBufPtr REC;         // shared with the operand batch (U and S[:,1:p] are part of the factorisation)
    ssm::DevBufPtr OUT;         // [ntiles][npo][nfo][32], record row: tail (tl) | Vnew (tr) | tau || pivot,d (tl+tm+1) | alpha (tp) | beta (tr) | A'U (tr)
    ssm::DevBufPtr GT;          // [ntiles][tr*tr][32]  U'U of each system (shared with the operand batch)
    ssm::DevBufPtr TT;          // [ntiles][tr][32]     U'b scratch of lmul!(Q', b)
};

namespace ssm {

constexpr int BPS_FRONT = 8;    // zero records in front of row 0 (the sweep looks tm <= 8 rows back)

// bps_kernels.cu
bool bps_pick_shape(int l, int m, int r, int p, int *tl, int *tm, int *tr, int *tp);
int bps_nfa(int tl, int tm, int tr, int tp);
int bps_nfo(int tl, int tm, int tr, int tp);
int launch_bps_pack(ssm_ctx *c, ssm_bps *A, const double *B, int64_t sb, const double *U, const double *V, int64_t suv,
                    const double *W, const double *S, int64_t sws);
int launch_bps_generate(ssm_ctx *c, ssm_bps *A, uint64_t seed);
int launch_bps_qr(ssm_ctx *c, const ssm_bps *A, ssm_bpsqr *F);
int launch_bps_gram(ssm_ctx *c, const ssm_bps *A);       // A->GT <- U'U (K5), marks it valid
int bps_own_rec(ssm_bps *A);                              // copy-on-write: records (and U'U) of its own before A's operands are rewritten
int launch_bpsqr_unpack(ssm_ctx *c, const ssm_bpsqr *F, double *B, int64_t sb, double *U, double *V, int64_t suv, double *W,
                        double *S, int64_t sws, double *tau, int64_t st);
int launch_bps_qt(ssm_ctx *c, const ssm_bpsqr *F, const double *rhs_in, double *rhs);
int launch_bps_backsolve(ssm_ctx *c, const ssm_bpsqr *F, const double *rhs_in, double *rhs);
int launch_bps_mul(ssm_ctx *c, const ssm_bps *A, const double *x, double *y, const double *b, double *relres);
int launch_bps_upper_ldiv(ssm_ctx *c, const ssm_bps *A, double *rhs);

// kernel-argument bundles (plain pointers; passed by value)
struct AbArgs {
    const double *AU; const double *V;      // operands
    double *RU; double *QT;                  // factors (QT may be null when reflectors are not kept)
    const double *rhs_in;                    // b (may be null; may alias rhs)
    double *rhs;                             // out: Q'b
    int n; int64_t np;
    int ncols;                               // reflectors formed: columns 0..ncols-1 (_almostbanded_qr!(A, tau, ncols), :137-139)
};
struct AbSolveArgs {
    const double *RU; const double *QT; const double *V;
    const double *rhs_in;                    // input vector batch (may alias rhs)
    double *rhs;                             // output / in-place vector batch
    int n; int64_t np; int unit;
};

// GroupRing geometry (stream_loader.cuh) for a grid of ntiles CTAs; false when the shared memory left per CTA is too small
bool pick_group(const ssm_ctx *c, size_t block_bytes, int64_t ntiles, int *nstage, int *gm);
// launchers implemented in ab_kernels.cu; return 0 or SSM_E_UNSUPPORTED when (l,u,r) has no fast instance
int launch_ab_qr(ssm_ctx *c, int l, int u, int r, int64_t ntiles, const AbArgs &a, bool with_rhs, bool store_q);
int launch_ab_qr_continue(ssm_ctx *c, int l, int u, int r, int64_t ntiles, const AbArgs &a, int kstart);
int launch_ab_grow_fix(ssm_ctx *c, int l, int u, int r, int64_t ntiles, const AbArgs &a, int n0, int k0);
int launch_ab_qt(ssm_ctx *c, int l, int u, int r, int64_t ntiles, const AbSolveArgs &a, bool transpose);
int launch_ab_backsolve(ssm_ctx *c, int l, int u, int r, int64_t ntiles, const AbSolveArgs &a);
// operand-level triangular solves (unfactored AlmostBandedMatrix), generic kernels only
int launch_ab_tri(ssm_ctx *c, const ssm_ab *A, double *rhs, bool upper, bool unit);

// util_kernels.cu
int launch_ab_pack(ssm_ctx *c, ssm_ab *A, const double *bands, int64_t sb, const double *U, int64_t su,
                   const double *V, int64_t sv, int64_t s0, int64_t count);
int launch_ab_pack_rows(ssm_ctx *c, ssm_ab *A, const double *bands, int64_t sb, int64_t jlo, const double *U, int64_t su, int64_t ldu,
                        const double *V, int64_t sv, int64_t i0u, int64_t s0, int64_t count, int64_t i0, int64_t i1);
int launch_ab_pack_vcat(ssm_ctx *c, ssm_ab *A, const double *a, int64_t sa, const double *Bd, int64_t sbd, int lb, int ub);
int launch_ab_unpack(ssm_ctx *c, const ssm_ab *A, double *bands, int64_t sb, double *U, int64_t su, double *V,
                     int64_t sv);
int launch_abqr_unpack(ssm_ctx *c, const ssm_abqr *F, double *factors, int64_t sf, double *U, int64_t su,
                       double *tau, int64_t st);
int launch_abqr_pack(ssm_ctx *c, ssm_abqr *F, const double *factors, int64_t sf, const double *U, int64_t su, const double *V, int64_t sv,
                     const double *tau, int64_t st);
int launch_ab_generate(ssm_ctx *c, ssm_ab *A, uint64_t seed, int dist);
int launch_vec_pack(ssm_ctx *c, ssm_vec *v, const double *b, int64_t stride, int64_t s0, int64_t count);
int launch_vec_unpack(ssm_ctx *c, const ssm_vec *v, double *b, int64_t stride, int64_t s0, int64_t count);
int launch_vec_generate(ssm_ctx *c, ssm_vec *v, uint64_t seed, int array_id);
int launch_ab_residual(ssm_ctx *c, const ssm_ab *A, const ssm_vec *x, const ssm_vec *b, double *relres_dev);

}  // namespace ssm
