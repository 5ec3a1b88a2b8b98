// bps_api.cu — C ABI entry points of the BandedPlusSemiseparable path (include/ssm_b200.h).
#include <algorithm>
#include "ssm_internal.cuh"

using namespace ssm;

namespace {
struct DevGuard {
    int prev = 0;
    explicit DevGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
    ~DevGuard() { int cur; cudaGetDevice(&cur); if (cur != prev) cudaSetDevice(prev); }
};
int chk_vec(const char *fn, int n, int64_t nb, const ssm_ctx *c, const ssm_vec *b) {
    if (!b) { set_error("%s: null vector", fn); return SSM_E_ARG; }
    if (b->n != n || b->nb != nb) { set_error("%s: DimensionMismatch: vector batch is %lld x %d, matrix batch is %lld x %d", fn, (long long)b->nb, b->n, (long long)nb, n); return SSM_E_SHAPE; }
    if (b->ctx->device != c->device) { set_error("%s: operands live on different devices", fn); return SSM_E_ARG; }
    uses(b->d, c);              // a vector of another context is used on this one's stream: its block leaves the free-list protocol
    return 0;
}
void bpsqr_uses(const ssm_bpsqr *F) { uses(F->REC, F->ctx); uses(F->OUT, F->ctx); uses(F->GT, F->ctx); uses(F->TT, F->ctx); }
// stage `count` doubles from a host or device pointer into the ctx staging area at offset `off`
int stage_in(ssm_ctx *c, double *stg, size_t off, const double *src, int64_t stride, size_t elems, int64_t nb, const double **out) {
    if (!src) { *out = nullptr; return 0; }
    if (is_device_ptr(src) && (size_t)stride == elems) { *out = src; return 0; }
    double *dst = stg + off;
    if ((size_t)stride == elems) SSM_CUDA(cudaMemcpyAsync(dst, src, elems * nb * sizeof(double), cudaMemcpyDefault, c->stream));
    else SSM_CUDA(cudaMemcpy2DAsync(dst, elems * sizeof(double), src, stride * sizeof(double), elems * sizeof(double), nb, cudaMemcpyDefault, c->stream));
    *out = dst;
    return 0;
}
}  // namespace

// A factor object shares the operand records (U and S[:, 1:p] are part of the factorisation): rewriting the operands of a batch that has
// live factor objects first gives the batch records of its own.
int ssm::bps_own_rec(ssm_bps *A) {
    if (!A->REC || A->REC.use_count() <= 1) return 0;
    DevBufPtr nr;
    SSM_TRY(dev_alloc(A->ctx, A->REC->count, true, nr));
    uses(A->REC, A->ctx);
    A->REC = nr;
    A->GT.reset(); A->gt_valid = false;           // the factor objects keep the old one
    return 0;
}

#define SSM_CHECK_ARG(cond, msg) do { if (!(cond)) { set_error("%s: %s", __func__, msg); return SSM_E_ARG; } } while (0)

extern "C" {

int ssm_bps_create(ssm_ctx *c, int n, int l, int m, int r, int p, int64_t nb, ssm_bps **out) {
    SSM_CHECK_ARG(c && out, "null argument");
    SSM_CHECK_ARG(n >= 1 && nb >= 1, "n and nb must be positive");
    SSM_CHECK_ARG(l >= 0 && m >= 0 && r >= 0 && p >= 0, "bandwidths and ranks must be non-negative");
    int tl, tm, tr, tp;
    if (!bps_pick_shape(l, m, r, p, &tl, &tm, &tr, &tp)) { set_error("BPS shape (l,m,r,p)=(%d,%d,%d,%d) exceeds the compiled kernel shapes (max (4,5,3,3))", l, m, r, p); return SSM_E_UNSUPPORTED; }
    DevGuard g(c->device);
    auto *A = new ssm_bps();
    A->ctx = c; A->ref.bind(c); A->n = n; A->l = l; A->m = m; A->r = r; A->p = p; A->tl = tl; A->tm = tm; A->tr = tr; A->tp = tp;
    A->nb = nb; A->ntiles = ntiles_of(nb); A->np = (int64_t)BPS_FRONT + n + PAD;
    int rc = dev_alloc(c, (size_t)(A->ntiles * A->np * bps_nfa(tl, tm, tr, tp) * 32), true, A->REC);
    if (rc) { delete A; return rc; }
    *out = A;
    return 0;
}
int ssm_bps_destroy(ssm_bps *A) { if (A) { obj_sync(A->ctx); delete A; } return 0; }

int ssm_bps_set(ssm_bps *A, const double *B, int64_t sb, const double *U, const double *V, int64_t suv, const double *W, const double *S, int64_t sws) {
    SSM_CHECK_ARG(A && B, "null argument");
    SSM_CHECK_ARG(A->r == 0 || (U && V), "null U/V");
    SSM_CHECK_ARG(A->p == 0 || (W && S), "null W/S");
    ssm_ctx *c = A->ctx;
    DevGuard g(c->device);
    const size_t eb = (size_t)(A->l + A->m + 1) * A->n, eu = (size_t)A->n * A->r, ew = (size_t)A->n * A->p;
    if (sb == 0) sb = (int64_t)eb;
    if (suv == 0) suv = (int64_t)eu;
    if (sws == 0) sws = (int64_t)ew;
    if ((size_t)sb < eb || (size_t)suv < eu || (size_t)sws < ew) { set_error("ssm_bps_set: Dimensions are not compatible."); return SSM_E_SHAPE; }
    SSM_TRY(bps_own_rec(A));
    double *stg;
    SSM_TRY(staging_dev(c, (size_t)A->nb * (eb + 2 * eu + 2 * ew) + 8, &stg));
    const double *dB, *dU, *dV, *dW, *dS;
    size_t off = 0;
    SSM_TRY(stage_in(c, stg, off, B, sb, eb, A->nb, &dB)); off += (size_t)A->nb * eb;
    SSM_TRY(stage_in(c, stg, off, U, suv, eu, A->nb, &dU)); off += (size_t)A->nb * eu;
    SSM_TRY(stage_in(c, stg, off, V, suv, eu, A->nb, &dV)); off += (size_t)A->nb * eu;
    SSM_TRY(stage_in(c, stg, off, W, sws, ew, A->nb, &dW)); off += (size_t)A->nb * ew;
    SSM_TRY(stage_in(c, stg, off, S, sws, ew, A->nb, &dS));
    // staged copies are dense; untouched device pointers keep the caller's stride (equal to dense by construction above)
    SSM_TRY(launch_bps_pack(c, A, dB, (int64_t)eb, dU, dV, (int64_t)eu, dW, dS, (int64_t)ew));
    SSM_TRY(launch_bps_gram(c, A));               // U'U goes with the operands (qr then reads the records once)
    SSM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int ssm_bps_generate(ssm_bps *A, uint64_t seed) {
    SSM_CHECK_ARG(A, "null argument");
    DevGuard g(A->ctx->device);
    SSM_TRY(bps_own_rec(A));
    SSM_TRY(launch_bps_generate(A->ctx, A, seed));
    return launch_bps_gram(A->ctx, A);
}

int ssm_bps_qr(const ssm_bps *A, ssm_bpsqr **out) {
    SSM_CHECK_ARG(A && out, "null argument");
    ssm_ctx *c = A->ctx;
    DevGuard g(c->device);
    ssm_bpsqr *F = *out;
    CrossOrder co(c, {F ? F->ctx : nullptr});
    if (F) {
        if (F->n != A->n || F->l != A->l || F->m != A->m || F->r != A->r || F->p != A->p || F->nb != A->nb) { set_error("ssm_bps_qr: factor object has a different shape"); return SSM_E_SHAPE; }
        F->REC = A->REC;
        uses(F->OUT, c); uses(F->TT, c);        // refactorisation of another context's factor object runs on A's stream
    } else {
        F = new ssm_bpsqr();
        F->ctx = c; F->ref.bind(c); F->n = A->n; F->l = A->l; F->m = A->m; F->r = A->r; F->p = A->p; F->tl = A->tl; F->tm = A->tm; F->tr = A->tr; F->tp = A->tp;
        F->nb = A->nb; F->ntiles = A->ntiles; F->np = A->np; F->npo = (int64_t)A->n + PAD;
        F->REC = A->REC;
        int rc = dev_alloc(c, (size_t)(F->ntiles * F->npo * bps_nfo(F->tl, F->tm, F->tr, F->tp) * 32), true, F->OUT);
        if (!rc) rc = dev_alloc(c, (size_t)(F->ntiles * F->tr * 32), true, F->TT);
        if (rc) { delete F; return rc; }
    }
    int rc = launch_bps_qr(c, A, F);
    if (rc) { if (!*out) delete F; return rc; }
    *out = F;
    return 0;
}
int ssm_bpsqr_destroy(ssm_bpsqr *F) { if (F) { obj_sync(F->ctx); delete F; } return 0; }

int ssm_bpsqr_get(const ssm_bpsqr *F, double *B, int64_t sb, double *U, double *V, int64_t suv, double *W, double *S, int64_t sws, double *tau, int64_t st) {
    SSM_CHECK_ARG(F, "null argument");
    ssm_ctx *c = F->ctx;
    DevGuard g(c->device);
    const size_t eb = (size_t)(2 * F->l + F->m + 1) * F->n, eu = (size_t)F->n * F->r, ew = (size_t)F->n * (F->p + F->r), et = (size_t)F->n;
    if (sb == 0) sb = (int64_t)eb;
    if (suv == 0) suv = (int64_t)eu;
    if (sws == 0) sws = (int64_t)ew;
    if (st == 0) st = (int64_t)et;
    // unpack into dense device staging, then copy each requested array to wherever the caller wants it
    double *stg;
    SSM_TRY(staging_dev(c, (size_t)F->nb * (eb + 2 * eu + 2 * ew + et) + 8, &stg));
    double *sB = stg, *sU = sB + (size_t)F->nb * eb, *sV = sU + (size_t)F->nb * eu, *sW = sV + (size_t)F->nb * eu, *sS = sW + (size_t)F->nb * ew, *sT = sS + (size_t)F->nb * ew;
    if (B) SSM_CUDA(cudaMemsetAsync(sB, 0, (size_t)F->nb * eb * sizeof(double), c->stream));
    SSM_TRY(launch_bpsqr_unpack(c, F, B ? sB : nullptr, (int64_t)eb, U ? sU : nullptr, V ? sV : nullptr, (int64_t)eu, W ? sW : nullptr, S ? sS : nullptr,
                                (int64_t)ew, tau ? sT : nullptr, (int64_t)et));
    auto outcopy = [&](double *dst, int64_t stride, const double *src, size_t elems) -> int {
        if (!dst || elems == 0) return 0;
        if ((size_t)stride == elems) SSM_CUDA(cudaMemcpyAsync(dst, src, elems * F->nb * sizeof(double), cudaMemcpyDefault, c->stream));
        else SSM_CUDA(cudaMemcpy2DAsync(dst, stride * sizeof(double), src, elems * sizeof(double), elems * sizeof(double), F->nb, cudaMemcpyDefault, c->stream));
        return 0;
    };
    SSM_TRY(outcopy(B, sb, sB, eb)); SSM_TRY(outcopy(U, suv, sU, eu)); SSM_TRY(outcopy(V, suv, sV, eu));
    SSM_TRY(outcopy(W, sws, sW, ew)); SSM_TRY(outcopy(S, sws, sS, ew)); SSM_TRY(outcopy(tau, st, sT, et));
    SSM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int ssm_bpsqr_lmul_qt(const ssm_bpsqr *F, ssm_vec *b) {
    SSM_CHECK_ARG(F, "null argument");
    SSM_TRY(chk_vec(__func__, F->n, F->nb, F->ctx, b));
    bpsqr_uses(F);
    DevGuard g(F->ctx->device);
    CrossOrder co(F->ctx, {b->ctx});
    return launch_bps_qt(F->ctx, F, b->d->p, b->d->p);
}
int ssm_bpsqr_upper_ldiv(const ssm_bpsqr *F, ssm_vec *b) {
    SSM_CHECK_ARG(F, "null argument");
    SSM_TRY(chk_vec(__func__, F->n, F->nb, F->ctx, b));
    bpsqr_uses(F);
    DevGuard g(F->ctx->device);
    CrossOrder co(F->ctx, {b->ctx});
    return launch_bps_backsolve(F->ctx, F, b->d->p, b->d->p);
}
int ssm_bpsqr_solve(const ssm_bpsqr *F, const ssm_vec *b, ssm_vec *x) {
    SSM_CHECK_ARG(F, "null argument");
    SSM_TRY(chk_vec(__func__, F->n, F->nb, F->ctx, b));
    SSM_TRY(chk_vec(__func__, F->n, F->nb, F->ctx, x));
    bpsqr_uses(F);
    DevGuard g(F->ctx->device);
    CrossOrder co(F->ctx, {b->ctx, x->ctx});
    SSM_TRY(launch_bps_qt(F->ctx, F, b->d->p, x->d->p));
    return launch_bps_backsolve(F->ctx, F, x->d->p, x->d->p);
}
int ssm_bpsqr_ldiv(const ssm_bpsqr *F, ssm_vec *b) { return ssm_bpsqr_solve(F, b, b); }

int ssm_bps_mul(const ssm_bps *A, const ssm_vec *x, ssm_vec *y) {
    SSM_CHECK_ARG(A, "null argument");
    SSM_TRY(chk_vec(__func__, A->n, A->nb, A->ctx, x));
    SSM_TRY(chk_vec(__func__, A->n, A->nb, A->ctx, y));
    if (x == y) { set_error("ssm_bps_mul: x and y must be distinct"); return SSM_E_ARG; }
    DevGuard g(A->ctx->device);
    CrossOrder co(A->ctx, {x->ctx, y->ctx});
    return launch_bps_mul(A->ctx, A, x->d->p, y->d->p, nullptr, nullptr);
}
int ssm_bps_residual(const ssm_bps *A, const ssm_vec *x, const ssm_vec *b, double *relres) {
    SSM_CHECK_ARG(A && relres, "null argument");
    SSM_TRY(chk_vec(__func__, A->n, A->nb, A->ctx, x));
    SSM_TRY(chk_vec(__func__, A->n, A->nb, A->ctx, b));
    ssm_ctx *c = A->ctx;
    DevGuard g(c->device);
    CrossOrder co(c, {x->ctx, b->ctx});
    if (is_device_ptr(relres)) return launch_bps_mul(c, A, x->d->p, nullptr, b->d->p, relres);
    double *stg;
    SSM_TRY(staging_dev(c, (size_t)A->ntiles * 32, &stg));
    SSM_TRY(launch_bps_mul(c, A, x->d->p, nullptr, b->d->p, stg));
    SSM_CUDA(cudaMemcpyAsync(relres, stg, (size_t)A->nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SSM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int ssm_bps_upper_ldiv(const ssm_bps *A, ssm_vec *b) {
    SSM_CHECK_ARG(A, "null argument");
    SSM_TRY(chk_vec(__func__, A->n, A->nb, A->ctx, b));
    DevGuard g(A->ctx->device);
    CrossOrder co(A->ctx, {b->ctx});
    return launch_bps_upper_ldiv(A->ctx, A, b->d->p);
}

}  // extern "C"
