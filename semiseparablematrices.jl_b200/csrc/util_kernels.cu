// util_kernels.cu — boundary kernels: reference layout <-> device record layout (K4), device-side synthetic
// operands (K11) and the structured residual (K9).  None of these is on the timed hot path; they are written
// for coalesced access on the device-layout side (lane fastest).
#include "../../include/ssm_rng.h"
#include "ssm_internal.cuh"

namespace ssm {

// ---- K4 pack: AlmostBanded -----------------------------------------------------------------------------
// thread = (tile, record i, lane); fields looped.  Row record of row i: A[i, i-l+d] = bands[(l+u-d) + (l+u+1)(i-l+d)]
// (band storage of AlmostBandedMatrix.jl:7-11), then U[i, q] = U[i + n q].  Lanes beyond nb become identity rows.
__global__ void ab_pack_kernel(double *AU, double *Vd, const double *bands, int64_t sb, const double *U, int64_t su,
                               const double *V, int64_t sv, int n, int l, int u, int r, int64_t nb, int64_t np,
                               int64_t s0, int64_t count) {
    const int lane = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    const int64_t tile = (int64_t)blockIdx.y + s0 / TILE;
    if (i >= n) return;
    const int64_t s = tile * TILE + lane;
    const bool live = s < nb && s >= s0 && s < s0 + count;
    if (!(s >= s0 && s < s0 + count) && s < nb) return;   // another chunk owns this lane
    const int nfa = l + u + 1 + r;
    double *rec = AU + ((tile * np + i) * nfa) * 32 + lane;
    const double *bs = bands + (s - s0) * sb;
    for (int d = 0; d <= l + u; ++d) {
        const int64_t j = i - l + d;
        double val = 0.0;
        if (live) { if (j >= 0 && j < n) val = bs[(int64_t)(l + u - d) + (int64_t)(l + u + 1) * j]; }
        else val = (d == l) ? 1.0 : 0.0;
        rec[d * 32] = val;
    }
    for (int q = 0; q < r; ++q) rec[(l + u + 1 + q) * 32] = live ? U[(s - s0) * su + i + (int64_t)n * q] : 0.0;
    double *vrec = Vd + ((tile * np + i) * (r > 0 ? r : 1)) * 32 + lane;
    for (int q = 0; q < r; ++q) vrec[q * 32] = live ? V[(s - s0) * sv + q + (int64_t)r * i] : 0.0;
}

// rows [i0, i1) of the records only (ssm_ab_set_rows).  The source arrays may be windows of the reference-layout arrays: band columns
// from jlo on (sb = doubles per system of the window), rows of U from i0u on with leading dimension ldu, columns of V from i0u on.
__global__ void ab_pack_rows_kernel(double *AU, double *Vd, const double *bands, int64_t sb, int64_t jlo, const double *U, int64_t su,
                                    int64_t ldu, const double *V, int64_t sv, int64_t i0u, int n, int l, int u, int r, int64_t nb,
                                    int64_t np, int64_t s0, int64_t count, int64_t i0, int64_t i1) {
    const int lane = threadIdx.x;
    const int64_t i = i0 + (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    const int64_t tile = (int64_t)blockIdx.y + s0 / TILE;
    const int64_t s = tile * TILE + lane;
    if (i >= i1) return;
    const bool live = s < nb && s >= s0 && s < s0 + count;
    if (!live && s < nb) return;                           // another chunk owns this lane; lanes past the batch hold identity rows
    const int nfa = l + u + 1 + r;
    double *rec = AU + ((tile * np + i) * nfa) * 32 + lane;
    const double *bs = bands + (s - s0) * sb;
    for (int d = 0; d <= l + u; ++d) {
        const int64_t j = i - l + d;
        double val = 0.0;
        if (live) { if (j >= 0 && j < n) val = bs[(int64_t)(l + u - d) + (int64_t)(l + u + 1) * (j - jlo)]; }
        else val = (d == l) ? 1.0 : 0.0;
        rec[d * 32] = val;
    }
    for (int q = 0; q < r; ++q) rec[(l + u + 1 + q) * 32] = live ? U[(s - s0) * su + (i - i0u) + ldu * q] : 0.0;
    double *vrec = Vd + ((tile * np + i) * (r > 0 ? r : 1)) * 32 + lane;
    for (int q = 0; q < r; ++q) vrec[q * 32] = live ? V[(s - s0) * sv + q + (int64_t)r * (i - i0u)] : 0.0;
}

__global__ void ab_unpack_kernel(const double *AU, const double *Vd, double *bands, int64_t sb, double *U, int64_t su,
                                 double *V, int64_t sv, int n, int l, int u, int r, int64_t nb, int64_t np) {
    const int lane = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    const int64_t tile = blockIdx.y;
    const int64_t s = tile * TILE + lane;
    if (i >= n || s >= nb) return;
    const int nfa = l + u + 1 + r;
    const double *rec = AU + ((tile * np + i) * nfa) * 32 + lane;
    if (bands) for (int d = 0; d <= l + u; ++d) {
        const int64_t j = i - l + d;
        if (j >= 0 && j < n) bands[s * sb + (int64_t)(l + u - d) + (int64_t)(l + u + 1) * j] = rec[d * 32];
    }
    if (U) for (int q = 0; q < r; ++q) U[s * su + i + (int64_t)n * q] = rec[(l + u + 1 + q) * 32];
    if (V) { const double *vrec = Vd + ((tile * np + i) * (r > 0 ? r : 1)) * 32 + lane;
             for (int q = 0; q < r; ++q) V[s * sv + q + (int64_t)r * i] = vrec[q * 32]; }
}

// ---- Vcat front end: A = Vcat(a, Bd), a dense r x n boundary rows, Bd an (n-r) x n BandedMatrix (lb, ub) ---------------------------
// What `_copyto!(::AlmostBandedLayout, ::VcatAlmostBandedLayout, dest, A)` (AlmostBandedMatrix.jl:316-337) builds on the host —
// U = [I_r; 0], V = a, bands[r+1:end, :] = Bd, in-band entries of the first r rows = a — written straight into the device
// records, so a spectral-method caller never materialises the (l+u+1) x n band array.  a: r x n column-major (= V's memory);
// Bd: (lb+ub+1) x n band storage, Bd[k, j] at row ub + k - j of column j.
__global__ void ab_pack_vcat_kernel(double *AU, double *Vd, const double *a, int64_t sa, const double *Bd, int64_t sbd, int lb, int ub,
                                    int n, int l, int u, int r, int64_t nb, int64_t np) {
    const int lane = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    const int64_t tile = blockIdx.y;
    if (i >= n) return;
    const int64_t s = tile * TILE + lane;
    const bool live = s < nb;
    const int nfa = l + u + 1 + r, ndb = lb + ub + 1;
    double *rec = AU + ((tile * np + i) * nfa) * 32 + lane;
    const double *as = a + s * sa, *bs = Bd + s * sbd;
    for (int d = 0; d <= l + u; ++d) {
        const int64_t j = i - l + d;
        double val = 0.0;
        if (!live) val = (d == l) ? 1.0 : 0.0;                    // padding lanes are identity systems
        else if (j >= 0 && j < n) {
            if (i < r) val = as[i + (int64_t)r * j];
            else { const int64_t k = i - r, off = j - k; if (off >= -lb && off <= ub) val = bs[(int64_t)(ub + k - j) + (int64_t)ndb * j]; }
        }
        rec[d * 32] = val;
    }
    for (int q = 0; q < r; ++q) rec[(l + u + 1 + q) * 32] = (live && i == q) ? 1.0 : 0.0;
    double *vrec = Vd + ((tile * np + i) * (r > 0 ? r : 1)) * 32 + lane;
    for (int q = 0; q < r; ++q) vrec[q * 32] = live ? as[q + (int64_t)r * i] : 0.0;
}

// zero the out-of-matrix corners of a band-storage batch (entries the loops above never write)
__global__ void band_corner_zero_kernel(double *bands, int64_t sb, int n, int lo, int up, int64_t nb) {
    const int64_t s = blockIdx.x;
    if (s >= nb) return;
    const int nd = lo + up + 1;
    for (int64_t e = threadIdx.x; e < (int64_t)nd * (up + lo); e += blockDim.x) {
        // columns 0..up-1 (top-left corner) and n-lo..n-1 (bottom-right corner)
        const int64_t jj = e / nd; const int d = (int)(e % nd);
        const int64_t j = jj < up ? jj : (int64_t)n - lo + (jj - up);
        if (j < 0 || j >= n) continue;
        const int64_t k = j + d - up;
        if (k < 0 || k >= n) bands[s * sb + d + (int64_t)nd * j] = 0.0;
    }
}

// QR factors -> reference layout: widened band storage (2l+u+1) x n (R in bands 0..ub, reflector tails in bands
// -1..-l), updated fill U (n x r), tau (n)   (AlmostBandedMatrix.jl:174-185)
__global__ void abqr_unpack_kernel(const double *RU, const double *QT, double *factors, int64_t sf, double *U, int64_t su,
                                   double *tau, int64_t st, int n, int l, int u, int r, int64_t nb, int64_t np) {
    const int lane = threadIdx.x;
    const int64_t k = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    const int64_t tile = blockIdx.y;
    const int64_t s = tile * TILE + lane;
    if (k >= n || s >= nb) return;
    const int ub = l + u, nfr = l + u + 1 + r, nfq = l + 1, nd = 2 * l + u + 1;
    const double *rrec = RU + ((tile * np + k) * nfr) * 32 + lane;
    const double *qrec = QT + ((tile * np + k) * nfq) * 32 + lane;
    if (factors) {
        for (int t = 0; t <= ub; ++t) { const int64_t j = k + t; if (j < n) factors[s * sf + (int64_t)(ub - t) + (int64_t)nd * j] = rrec[t * 32]; }
        for (int i = 1; i <= l; ++i) if (k + i < n) factors[s * sf + (int64_t)(ub + i) + (int64_t)nd * k] = qrec[(i - 1) * 32];
    }
    if (U) for (int q = 0; q < r; ++q) U[s * su + k + (int64_t)n * q] = rrec[(ub + 1 + q) * 32];
    if (tau) tau[s * st + k] = qrec[l * 32];
}

// reference layout -> QR factor records (the inverse of abqr_unpack_kernel, plus the V records): what a factor object that was not made by this
// library instance (a copy, a deserialised QR) is re-created from.  Entries outside the matrix stay zero (the buffers are zero-initialised).
__global__ void abqr_pack_kernel(double *RU, double *QT, double *Vr, const double *factors, int64_t sf, const double *U, int64_t su, const double *V,
                                 int64_t sv, const double *tau, int64_t st, int n, int l, int u, int r, int64_t nb, int64_t np) {
    const int lane = threadIdx.x;
    const int64_t k = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    const int64_t tile = blockIdx.y;
    const int64_t s = tile * TILE + lane;
    if (k >= n) return;
    const bool live = s < nb;                       // spare lanes of the last tile hold identity rows: nothing divides by zero
    const int ub = l + u, nfr = l + u + 1 + r, nfq = l + 1, nd = 2 * l + u + 1;
    double *rrec = RU + ((tile * np + k) * nfr) * 32 + lane;
    double *qrec = QT + ((tile * np + k) * nfq) * 32 + lane;
    for (int t = 0; t <= ub; ++t) { const int64_t j = k + t; rrec[t * 32] = live ? (j < n ? factors[s * sf + (int64_t)(ub - t) + (int64_t)nd * j] : 0.0) : (t == 0 ? 1.0 : 0.0); }
    for (int i = 1; i <= l; ++i) qrec[(i - 1) * 32] = (live && k + i < n) ? factors[s * sf + (int64_t)(ub + i) + (int64_t)nd * k] : 0.0;
    for (int q = 0; q < r; ++q) {
        rrec[(ub + 1 + q) * 32] = live ? U[s * su + k + (int64_t)n * q] : 0.0;
        Vr[((tile * np + k) * r + q) * 32 + lane] = live ? V[s * sv + q + (int64_t)r * k] : 0.0;
    }
    qrec[l * 32] = live ? tau[s * st + k] : 0.0;
}

// ---- vectors: [nb][n] <-> [tile][rec][32] through a 32x32 shared-memory transpose -------------------------
__global__ void vec_pack_kernel(double *d, const double *b, int64_t stride, int n, int64_t nb, int64_t np, int64_t s0, int64_t count) {
    __shared__ double t[32][33];
    const int64_t tile = (int64_t)blockIdx.y + s0 / TILE;
    const int64_t i0 = (int64_t)blockIdx.x * 32;
    for (int sl = threadIdx.y; sl < 32; sl += blockDim.y) {
        const int64_t s = tile * TILE + sl, i = i0 + threadIdx.x;
        t[sl][threadIdx.x] = (s >= s0 && s < s0 + count && s < nb && i < n) ? b[(s - s0) * stride + i] : 0.0;
    }
    __syncthreads();
    for (int il = threadIdx.y; il < 32; il += blockDim.y) {
        const int64_t i = i0 + il, s = tile * TILE + threadIdx.x;
        const bool mine = (s >= s0 && s < s0 + count) || s >= nb;
        if (i < n && mine) d[(tile * np + i) * 32 + threadIdx.x] = t[threadIdx.x][il];
    }
}
__global__ void vec_unpack_kernel(const double *d, double *b, int64_t stride, int n, int64_t nb, int64_t np, int64_t s0, int64_t count) {
    __shared__ double t[32][33];
    const int64_t tile = (int64_t)blockIdx.y + s0 / TILE;
    const int64_t i0 = (int64_t)blockIdx.x * 32;
    for (int il = threadIdx.y; il < 32; il += blockDim.y) {
        const int64_t i = i0 + il;
        t[il][threadIdx.x] = (i < n) ? d[(tile * np + i) * 32 + threadIdx.x] : 0.0;
    }
    __syncthreads();
    for (int sl = threadIdx.y; sl < 32; sl += blockDim.y) {
        const int64_t s = tile * TILE + sl, i = i0 + threadIdx.x;
        if (s >= s0 && s < s0 + count && s < nb && i < n) b[(s - s0) * stride + i] = t[threadIdx.x][sl];
    }
}

// ---- K11 synthetic operands, written straight into the device layout ---------------------------------------
__global__ void ab_generate_kernel(double *AU, double *Vd, int n, int l, int u, int r, int64_t nb, int64_t np, uint64_t seed, int dist) {
    const int lane = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    const int64_t tile = blockIdx.y;
    if (i >= n) return;
    const int64_t s = tile * TILE + lane;
    const bool live = s < nb;
    const uint64_t key = ssm_syskey(seed, (uint64_t)s);
    const int nfa = l + u + 1 + r;
    double *rec = AU + ((tile * np + i) * nfa) * 32 + lane;
    for (int d = 0; d <= l + u; ++d) {
        const int64_t j = i - l + d;        // column of this entry; band-storage row of (i,j) is u + i - j = l + u - d
        double val = (d == l) ? 1.0 : 0.0;
        if (live) val = (j >= 0 && j < n) ? ssm_ab_band_elem(key, dist, n, l, u, r, l + u - d, (int)j) : 0.0;
        rec[d * 32] = val;
    }
    for (int q = 0; q < r; ++q) rec[(l + u + 1 + q) * 32] = live ? ssm_ab_U_elem(key, dist, n, r, (int)i, q) : 0.0;
    double *vrec = Vd + ((tile * np + i) * (r > 0 ? r : 1)) * 32 + lane;
    for (int q = 0; q < r; ++q) vrec[q * 32] = live ? ssm_ab_V_elem(key, dist, n, r, q, (int)i) : 0.0;
}
__global__ void vec_generate_kernel(double *d, int n, int64_t nb, int64_t np, uint64_t seed, int array_id) {
    const int lane = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    const int64_t tile = blockIdx.y;
    if (i >= n) return;
    const int64_t s = tile * TILE + lane;
    d[(tile * np + i) * 32 + lane] = (s < nb) ? 2.0 * ssm_unif_key(ssm_syskey(seed, (uint64_t)s), (uint64_t)array_id, (uint64_t)i) - 1.0 : 0.0;
}

// ---- K9 residual: relres[s] = ||b - A x|| / ||b||, A via getindex semantics (AlmostBandedMatrix.jl:84-90) -----
__global__ void __launch_bounds__(32) ab_residual_kernel(const double *AU, const double *Vd, const double *x_, const double *b_,
                                                         double *relres, int n, int l, int u, int r, int64_t nb, int64_t np) {
    const int lane = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t s = tile * TILE + lane;
    const int nfa = l + u + 1 + r;
    const double *au = AU + tile * np * nfa * 32 + lane;
    const double *vv = Vd + tile * np * (r > 0 ? r : 1) * 32 + lane;
    const double *x = x_ + tile * np * 32 + lane;
    const double *b = b_ + tile * np * 32 + lane;
    double S[SSM_MAX_RANK];
    for (int q = 0; q < SSM_MAX_RANK; ++q) S[q] = 0.0;
    double num = 0.0, den = 0.0;
    for (int64_t i = n - 1; i >= 0; --i) {
        const int64_t c = i + u + 1;                      // first fill column of row i
        if (c < n) { const double xc = x[c * 32]; for (int q = 0; q < r; ++q) S[q] = fma(vv[(c * r + q) * 32], xc, S[q]); }
        double ax = 0.0;
        for (int d = 0; d <= l + u; ++d) { const int64_t j = i - l + d; if (j >= 0 && j < n) ax = fma(au[(i * nfa + d) * 32], x[j * 32], ax); }
        for (int q = 0; q < r; ++q) ax = fma(au[(i * nfa + l + u + 1 + q) * 32], S[q], ax);
        const double bi = b[i * 32], dlt = bi - ax;
        num = fma(dlt, dlt, num); den = fma(bi, bi, den);
    }
    if (s < nb) relres[s] = sqrt(num) / sqrt(den);
}

// ---- launch wrappers -------------------------------------------------------------------------------------
static dim3 rec_grid(int n, int64_t ntiles) { return dim3((unsigned)((n + 7) / 8), (unsigned)ntiles); }

int launch_ab_pack(ssm_ctx *c, ssm_ab *A, const double *bands, int64_t sb, const double *U, int64_t su, const double *V,
                   int64_t sv, int64_t s0, int64_t count) {
    if (count <= 0) return 0;
    if (s0 % TILE != 0) { set_error("pack chunk must start on a tile boundary"); return SSM_E_ARG; }
    int64_t send = s0 + count;
    int64_t tiles = (send + TILE - 1) / TILE - s0 / TILE;
    ab_pack_kernel<<<rec_grid(A->n, tiles), dim3(32, 8), 0, c->stream>>>(A->AU->p, A->V->p, bands, sb, U, su, V, sv, A->n, A->l, A->u,
                                                                          A->r, A->nb, A->np, s0, count);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}
int launch_ab_pack_rows(ssm_ctx *c, ssm_ab *A, const double *bands, int64_t sb, int64_t jlo, const double *U, int64_t su, int64_t ldu,
                        const double *V, int64_t sv, int64_t i0u, int64_t s0, int64_t count, int64_t i0, int64_t i1) {
    if (count <= 0 || i1 <= i0) return 0;
    if (s0 % TILE != 0) { set_error("pack chunk must start on a tile boundary"); return SSM_E_ARG; }
    const int64_t tiles = (s0 + count + TILE - 1) / TILE - s0 / TILE;
    ab_pack_rows_kernel<<<rec_grid((int)(i1 - i0), tiles), dim3(32, 8), 0, c->stream>>>(A->AU->p, A->V->p, bands, sb, jlo, U, su, ldu, V, sv, i0u, A->n, A->l,
                                                                                         A->u, A->r, A->nb, A->np, s0, count, i0, i1);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}
int launch_ab_pack_vcat(ssm_ctx *c, ssm_ab *A, const double *a, int64_t sa, const double *Bd, int64_t sbd, int lb, int ub) {
    ab_pack_vcat_kernel<<<rec_grid(A->n, A->ntiles), dim3(32, 8), 0, c->stream>>>(A->AU->p, A->V->p, a, sa, Bd, sbd, lb, ub, A->n, A->l, A->u, A->r, A->nb, A->np);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}
int launch_ab_unpack(ssm_ctx *c, const ssm_ab *A, double *bands, int64_t sb, double *U, int64_t su, double *V, int64_t sv) {
    if (bands) { band_corner_zero_kernel<<<(unsigned)A->nb, 64, 0, c->stream>>>(bands, sb, A->n, A->l, A->u, A->nb); c->launches++; }
    ab_unpack_kernel<<<rec_grid(A->n, A->ntiles), dim3(32, 8), 0, c->stream>>>(A->AU->p, A->V->p, bands, sb, U, su, V, sv, A->n, A->l,
                                                                                A->u, A->r, A->nb, A->np);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}
int launch_abqr_unpack(ssm_ctx *c, const ssm_abqr *F, double *factors, int64_t sf, double *U, int64_t su, double *tau, int64_t st) {
    if (factors) { band_corner_zero_kernel<<<(unsigned)F->nb, 64, 0, c->stream>>>(factors, sf, F->n, F->l, F->l + F->u, F->nb); c->launches++; }
    abqr_unpack_kernel<<<rec_grid(F->n, F->ntiles), dim3(32, 8), 0, c->stream>>>(F->RU->p, F->QT->p, factors, sf, U, su, tau, st, F->n,
                                                                                  F->l, F->u, F->r, F->nb, F->np);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}
int launch_abqr_pack(ssm_ctx *c, ssm_abqr *F, const double *factors, int64_t sf, const double *U, int64_t su, const double *V, int64_t sv,
                     const double *tau, int64_t st) {
    abqr_pack_kernel<<<rec_grid(F->n, F->ntiles), dim3(32, 8), 0, c->stream>>>(F->RU->p, F->QT->p, F->V->p, factors, sf, U, su, V, sv, tau, st, F->n, F->l,
                                                                                F->u, F->r, F->nb, F->np);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}
int launch_ab_generate(ssm_ctx *c, ssm_ab *A, uint64_t seed, int dist) {
    ab_generate_kernel<<<rec_grid(A->n, A->ntiles), dim3(32, 8), 0, c->stream>>>(A->AU->p, A->V->p, A->n, A->l, A->u, A->r, A->nb, A->np, seed, dist);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}
int launch_vec_pack(ssm_ctx *c, ssm_vec *v, const double *b, int64_t stride, int64_t s0, int64_t count) {
    if (count <= 0) return 0;
    if (s0 % TILE != 0) { set_error("pack chunk must start on a tile boundary"); return SSM_E_ARG; }
    int64_t tiles = (s0 + count + TILE - 1) / TILE - s0 / TILE;
    vec_pack_kernel<<<dim3((unsigned)((v->n + 31) / 32), (unsigned)tiles), dim3(32, 8), 0, c->stream>>>(v->d->p, b, stride, v->n, v->nb, v->np, s0, count);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}
int launch_vec_unpack(ssm_ctx *c, const ssm_vec *v, double *b, int64_t stride, int64_t s0, int64_t count) {
    if (count <= 0) return 0;
    int64_t tiles = (s0 + count + TILE - 1) / TILE - s0 / TILE;
    vec_unpack_kernel<<<dim3((unsigned)((v->n + 31) / 32), (unsigned)tiles), dim3(32, 8), 0, c->stream>>>(v->d->p, b, stride, v->n, v->nb, v->np, s0, count);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}
int launch_vec_generate(ssm_ctx *c, ssm_vec *v, uint64_t seed, int array_id) {
    vec_generate_kernel<<<rec_grid(v->n, v->ntiles), dim3(32, 8), 0, c->stream>>>(v->d->p, v->n, v->nb, v->np, seed, array_id);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}
int launch_ab_residual(ssm_ctx *c, const ssm_ab *A, const ssm_vec *x, const ssm_vec *b, double *relres_dev) {
    ab_residual_kernel<<<(unsigned)A->ntiles, 32, 0, c->stream>>>(A->AU->p, A->V->p, x->d->p, b->d->p, relres_dev, A->n, A->l, A->u, A->r, A->nb, A->np);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ssm
