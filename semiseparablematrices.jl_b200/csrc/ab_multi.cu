// ab_multi.cu — ldiv!(F::QR{<:Any,<:AlmostBandedMatrix}, B::AbstractMatrix) (AlmostBandedMatrix.jl:194-199): several right-hand
// sides per system in ONE pass over the factors.  A single-vector ldiv! moves 16n doubles per system of which only 3n belong to
// the vector; here the reflectors (K2) and the R rows (K3) are read once for a group of W = 4 or 8 columns, whose W x (l+1)
// (resp. W x (ub+1+r)) states live in registers.  Same per-system arithmetic as the single-vector kernels (ab_core.cuh: qt_step,
// bs_step), so column j of the result is bit-identical to ldiv!(F, B[:, j]).
//
// Device layout of a matrix batch: [tile][group][row][W][32] — inside a group a row record is W fields x 32 lanes, so the
// stream loaders of stream_loader.cuh pull it exactly like a vector.  nrhs is padded with zero columns to a multiple of W.
#include <algorithm>
#include "ab_core.cuh"
#include "ssm_internal.cuh"
#include "stream_loader.cuh"

struct ssm_mat {
    ssm_ctx *ctx; ssm::CtxRef ref; int n, nrhs, w, ng; int64_t nb, ntiles, np;
    ssm::DevBufPtr d;           // [ntiles][ng][np][w][32]
};

namespace ssm {

// ---- pack / unpack: per system the Julia Matrix memory (n x nrhs column-major), systems `stride` doubles apart ------------------
__global__ void mat_pack_kernel(double *d, const double *B, int64_t stride, int n, int nrhs, int w, int ng, int64_t nb, int64_t np) {
    __shared__ double t[32][33];
    const int64_t tile = blockIdx.y;
    const int j = blockIdx.z;                              // column (padded columns are zero from the allocation)
    const int64_t i0 = (int64_t)blockIdx.x * 32;
    for (int sl = threadIdx.y; sl < 32; sl += blockDim.y) {
        const int64_t s = tile * TILE + sl, i = i0 + threadIdx.x;
        t[sl][threadIdx.x] = (s < nb && i < n) ? B[s * stride + (int64_t)j * n + i] : 0.0;
    }
    __syncthreads();
    const int g = j / w, jj = j % w;
    for (int il = threadIdx.y; il < 32; il += blockDim.y) {
        const int64_t i = i0 + il;
        if (i < n) d[((((tile * ng + g) * np + i) * w) + jj) * 32 + threadIdx.x] = t[threadIdx.x][il];
    }
}
__global__ void mat_unpack_kernel(const double *d, double *B, int64_t stride, int n, int nrhs, int w, int ng, int64_t nb, int64_t np) {
    __shared__ double t[32][33];
    const int64_t tile = blockIdx.y;
    const int j = blockIdx.z;
    const int64_t i0 = (int64_t)blockIdx.x * 32;
    const int g = j / w, jj = j % w;
    for (int il = threadIdx.y; il < 32; il += blockDim.y) {
        const int64_t i = i0 + il;
        t[threadIdx.x][il] = (i < n) ? d[((((tile * ng + g) * np + i) * w) + jj) * 32 + threadIdx.x] : 0.0;
    }
    __syncthreads();
    for (int sl = threadIdx.y; sl < 32; sl += blockDim.y) {
        const int64_t s = tile * TILE + sl, i = i0 + threadIdx.x;
        if (s < nb && i < n) B[s * stride + (int64_t)j * n + i] = t[sl][threadIdx.x];
    }
}
// column j of a matrix batch <-> a vector batch (generic-shape fallback: one single-vector solve per column)
__global__ void mat_col_copy_kernel(double *mat, double *vec, int n, int w, int ng, int64_t np, int j, int to_vec) {
    const int lane = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y, tile = blockIdx.y;
    if (i >= n) return;
    const int g = j / w, jj = j % w;
    double *m = mat + ((((tile * ng + g) * np + i) * w) + jj) * 32 + lane, *v = vec + (tile * np + i) * 32 + lane;
    if (to_vec) *v = *m; else *m = *v;
}

// ---- K2 for W columns: B <- Q'B --------------------------------------------------------------------------------------------
template <int L, int W>
__global__ void __launch_bounds__(32) ab_qt_multi_kernel(const double *QT, double *B, int n, int64_t np, int ng, int nstage) {
    constexpr int P = L + 1, NFQ = L + 1;
    extern __shared__ __align__(128) unsigned char smem[];
    const int64_t tile = blockIdx.x / ng;
    const int lane = threadIdx.x;
    const double *qt = QT + tile * np * NFQ * 32 + lane;
    double *bb = B + (int64_t)blockIdx.x * np * W * 32 + lane;          // blockIdx.x = tile * ng + group
    QtState<L> st[W];
#pragma unroll
    for (int j = 0; j < W; ++j) qt_prologue<L>(st[j], [&](int rho) { return bb[(rho * W + j) * 32]; });
    const int64_t nsteps = ((n + P - 1) / P) * P;
    RingLoader<NFQ, W, 0, +1> ld;
    ld.init(qt, bb + (int64_t)L * W * 32, nullptr, smem, nstage, nsteps);
    for (int kb = 0; kb < n; kb += P) {
#pragma unroll
        for (int ph = 0; ph < P; ++ph) {
            const int k = kb + ph;
            double q[NFQ], bn[W], dummy[1];
            ld.acquire(k, q, bn, dummy);
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const double c = qt_step<L>(st[j], ph, q, bn[j]);
                if (k < n) bb[((int64_t)k * W + j) * 32] = c;
            }
        }
    }
}

// ---- K3 for W columns: B <- R \ B ------------------------------------------------------------------------------------------
template <int L, int U, int R, int W>
__global__ void __launch_bounds__(32) ab_bs_multi_kernel(const double *RU, const double *V, double *B, int n, int64_t np, int ng, int nstage) {
    using D = AbDims<L, U, R>;
    constexpr int UB = D::UB, PX = UB + 1, NFR = D::NFR;
    extern __shared__ __align__(128) unsigned char smem[];
    const int64_t tile = blockIdx.x / ng;
    const int lane = threadIdx.x;
    const double *ru = RU + tile * np * NFR * 32 + lane;
    const double *vv = V + tile * np * (R > 0 ? R : 1) * 32 + lane;
    double *bb = B + (int64_t)blockIdx.x * np * W * 32 + lane;
    BsState<UB, R> st[W];
#pragma unroll
    for (int j = 0; j < W; ++j) bs_init<UB, R>(st[j]);
    const int nblk = (n + PX - 1) / PX;
    const int itop = nblk * PX - 1;
    const int64_t nsteps = (int64_t)nblk * PX;
    RingLoader<NFR, R, W, -1> ld;
    ld.init(ru + (int64_t)itop * NFR * 32, vv + (int64_t)(itop + UB + 1) * R * 32, bb + (int64_t)itop * W * 32, smem, nstage, nsteps);
    int64_t gstep = 0;
    for (int ib = (nblk - 1) * PX; ib >= 0; ib -= PX) {
#pragma unroll
        for (int ph = PX - 1; ph >= 0; --ph) {
            const int i = ib + ph;
            double r[NFR], vc[R > 0 ? R : 1], cc[W];
            ld.acquire(gstep++, r, vc, cc);
            if (i < n) {
#pragma unroll
                for (int j = 0; j < W; ++j) bb[((int64_t)i * W + j) * 32] = bs_step<UB, R>(st[j], ph, r, vc, cc[j]);
            }
        }
    }
}

static int stages_for(size_t stage_bytes) {
    int s = (int)(20000 / (stage_bytes + 8));
    return std::min(8, std::max(2, s));
}

template <int L, int U, int R, int W>
static int multi_inst(ssm_ctx *c, const ssm_abqr *F, ssm_mat *B) {
    using D = AbDims<L, U, R>;
    const unsigned grid = (unsigned)(F->ntiles * B->ng);
    {
        using LD = RingLoader<L + 1, W, 0, +1>;
        const int ns = stages_for(LD::STAGE_BYTES);
        const size_t smem = LD::smem_bytes(ns);
        auto kern = ab_qt_multi_kernel<L, W>;
        if (smem > 48 * 1024) SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 32, smem, c->stream>>>(F->QT->p, B->d->p, F->n, F->np, B->ng, ns);
    }
    {
        using LD = RingLoader<D::NFR, R, W, -1>;
        const int ns = stages_for(LD::STAGE_BYTES);
        const size_t smem = LD::smem_bytes(ns);
        auto kern = ab_bs_multi_kernel<L, U, R, W>;
        if (smem > 48 * 1024) SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 32, smem, c->stream>>>(F->RU->p, F->V->p, B->d->p, F->n, F->np, B->ng, ns);
    }
    c->launches += 2;
    SSM_CUDA(cudaGetLastError());
    return 0;
}

#define SSM_ABM_SHAPES(X) X(1, 0, 1) X(1, 1, 1) X(1, 0, 2) X(2, 0, 2) X(2, 1, 2) X(2, 2, 1) X(2, 2, 2) X(3, 3, 2) X(3, 3, 3) X(4, 4, 2)

}  // namespace ssm

using namespace ssm;
#define SSM_CHECK_ARG(cond, msg) do { if (!(cond)) { set_error("%s: %s", __func__, msg); return SSM_E_ARG; } } while (0)

extern "C" {

int ssm_mat_create(ssm_ctx *c, int n, int nrhs, int64_t nb, ssm_mat **out) {
    SSM_CHECK_ARG(c && out, "null argument");
    SSM_CHECK_ARG(n >= 1 && nrhs >= 1 && nb >= 1, "sizes must be positive");
    auto *B = new ssm_mat();
    B->ctx = c; B->ref.bind(c); B->n = n; B->nrhs = nrhs; B->nb = nb; B->ntiles = ntiles_of(nb); B->np = (int64_t)n + PAD;
    B->w = nrhs <= 4 ? 4 : 8;
    B->ng = (nrhs + B->w - 1) / B->w;
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(c->device);
    int rc = dev_alloc(c, (size_t)(B->ntiles * B->ng * B->np * B->w * 32), true, B->d);
    cudaSetDevice(prev);
    if (rc) { delete B; return rc; }
    *out = B;
    return 0;
}
int ssm_mat_destroy(ssm_mat *B) { if (B) { obj_sync(B->ctx); delete B; } return 0; }

static int mat_xfer(ssm_mat *B, double *host_or_dev, int64_t stride, bool set) {
    ssm_ctx *c = B->ctx;
    const size_t per = (size_t)B->n * B->nrhs;
    if (stride == 0) stride = (int64_t)per;
    if ((size_t)stride < per) { set_error("ssm_mat: stride smaller than one n x nrhs matrix"); return SSM_E_SHAPE; }
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(c->device);
    int rc = 0;
    const bool dev = is_device_ptr(host_or_dev);
    double *p = host_or_dev;
    int64_t st = stride;
    if (!dev) {
        rc = staging_dev(c, (size_t)B->nb * per, &p);
        st = (int64_t)per;
        if (!rc && set) {
            cudaError_t e = (size_t)stride == per ? cudaMemcpyAsync(p, host_or_dev, B->nb * per * sizeof(double), cudaMemcpyHostToDevice, c->stream)
                                                  : cudaMemcpy2DAsync(p, per * sizeof(double), host_or_dev, stride * sizeof(double), per * sizeof(double), B->nb, cudaMemcpyHostToDevice, c->stream);
            if (e != cudaSuccess) rc = cuda_fail(e, "H2D matrix batch", __FILE__, __LINE__);
        }
    }
    if (!rc) {
        dim3 grid((unsigned)((B->n + 31) / 32), (unsigned)B->ntiles, (unsigned)B->nrhs), block(32, 8);
        if (set) mat_pack_kernel<<<grid, block, 0, c->stream>>>(B->d->p, p, st, B->n, B->nrhs, B->w, B->ng, B->nb, B->np);
        else mat_unpack_kernel<<<grid, block, 0, c->stream>>>(B->d->p, p, st, B->n, B->nrhs, B->w, B->ng, B->nb, B->np);
        c->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = cuda_fail(e, "matrix pack/unpack", __FILE__, __LINE__);
    }
    if (!rc && !dev) {
        cudaError_t e = cudaSuccess;
        if (!set) e = (size_t)stride == per ? cudaMemcpyAsync(host_or_dev, p, B->nb * per * sizeof(double), cudaMemcpyDeviceToHost, c->stream)
                                            : cudaMemcpy2DAsync(host_or_dev, stride * sizeof(double), p, per * sizeof(double), per * sizeof(double), B->nb, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "matrix batch copy", __FILE__, __LINE__);
    }
    cudaSetDevice(prev);
    return rc;
}
int ssm_mat_set(ssm_mat *B, const double *data, int64_t stride) {
    SSM_CHECK_ARG(B && data, "null argument");
    return mat_xfer(B, const_cast<double *>(data), stride, true);
}
int ssm_mat_get(const ssm_mat *B, double *data, int64_t stride) {
    SSM_CHECK_ARG(B && data, "null argument");
    return mat_xfer(const_cast<ssm_mat *>(B), data, stride, false);
}

// ldiv!(F, B::AbstractMatrix) (AlmostBandedMatrix.jl:194-199): B <- R \ (Q'B), in place
int ssm_abqr_ldiv_mat(const ssm_abqr *F, ssm_mat *B) {
    SSM_CHECK_ARG(F && B, "null argument");
    if (B->n != F->n || B->nb != F->nb) { set_error("ssm_abqr_ldiv_mat: DimensionMismatch: matrix batch is %lld x (%d x %d), factor batch is %lld x %d", (long long)B->nb, B->n, B->nrhs, (long long)F->nb, F->n); return SSM_E_SHAPE; }
    if (!F->QT) { set_error("factor object was created without reflectors"); return SSM_E_STATE; }
    if (F->ncols < F->n - 1) { set_error("ssm_abqr_ldiv_mat: the factor object holds %d of %d reflectors: finish it first (ssm_abqr_continue)", F->ncols, F->n - 1); return SSM_E_STATE; }
    ssm_ctx *c = F->ctx;
    uses(B->d, c); uses(F->V, c); uses(F->RU, c); uses(F->QT, c);
    CrossOrder co(c, {B->ctx});
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(c->device);
    int rc = SSM_E_UNSUPPORTED;
    if (c->ab_variant != 9) {
#define X(L_, U_, R_)                                                                                                      \
    if (F->l == L_ && F->u == U_ && F->r == R_) rc = B->w == 4 ? multi_inst<L_, U_, R_, 4>(c, F, B) : multi_inst<L_, U_, R_, 8>(c, F, B);
        SSM_ABM_SHAPES(X)
#undef X
    }
    if (rc == SSM_E_UNSUPPORTED) {
        // shapes without a register-resident instance: one single-vector ldiv! (generic kernels) per column
        ssm_vec *v = nullptr;
        rc = ssm_vec_create(c, F->n, F->nb, &v);
        for (int j = 0; j < B->nrhs && !rc; ++j) {
            dim3 grid((unsigned)((F->n + 7) / 8), (unsigned)F->ntiles), block(32, 8);
            mat_col_copy_kernel<<<grid, block, 0, c->stream>>>(B->d->p, v->d->p, F->n, B->w, B->ng, F->np, j, 1);
            rc = ssm_abqr_ldiv(F, v);
            mat_col_copy_kernel<<<grid, block, 0, c->stream>>>(B->d->p, v->d->p, F->n, B->w, B->ng, F->np, j, 0);
            c->launches += 2;
        }
        if (v) ssm_vec_destroy(v);
    }
    cudaSetDevice(prev);
    return rc;
}

}  // extern "C"
