/* ssm_b200.h — C ABI of libssmb200.so: the B200-native (sm_100a) batched structured-QR hot path of
 * SemiseparableMatrices.jl.  Plain pointers and sizes only; no torch / CUDA types in any signature.
 *
 * Each entry point names the reference interface it replaces (file:line in /root/reference/src).  The
 * reference has no batch API — a batch is a Vector of its matrices — so every call takes `nb` systems of
 * identical shape; nb = 1 is the single-matrix call the Julia method overrides make.
 *
 * Memory convention at the boundary ("reference layout", 8-byte aligned, per system, then batch-strided):
 *   AlmostBanded   bands : (l+u+1) x n column-major band storage, A[k,j] at row u+1+k-j of column j
 *                          (BandedMatrices `.data`; AlmostBandedMatrix.jl:7-11)
 *                  U     : n x r column-major, V : r x n column-major  (`A.fill.args`, SemiseparableMatrices.jl:20-26)
 *   BPS            B     : (l+m+1) x n band storage; U,V : n x r; W,S : n x p  (BandedPlusSemiseparableMatrix.jl:1-8)
 *   vectors        b     : n doubles
 *   system s of a batch starts at ptr + s*stride (stride in doubles; 0 selects the dense packing).
 * Pointers may be HOST or DEVICE addresses; the library detects which (cudaPointerGetAttributes).  Host
 * pointers are staged through pinned memory; the library never retains a caller pointer after return.
 *
 * Inside the library the operands live in a batch-interleaved, tile-blocked layout in HBM.
 *
 * Return value: 0 ok; < 0 argument / shape error (SSM_E_*; the Julia glue maps them to DimensionMismatch /
 * ArgumentError as the reference throws at AlmostBandedMatrix.jl:14-16,310 and semiseparableqr.jl:46,105-107);
 * > 0 CUDA runtime failure (cudaError_t).  ssm_last_error() returns the message of the calling thread's
 * last failure.  No singularity checks are made — like the reference, Inf/NaN propagate.
 *
 * Threading: handles are not shared-state; calls on different ssm_ctx from different host threads are
 * safe.  Every call enqueues on its ctx's stream and returns after the work is ENQUEUED unless it copies
 * to a host pointer (then it synchronises); use ssm_ctx_sync() to wait.  An entry point given objects of
 * ANOTHER context of the same device (a vector of ctx 2 solved with a factorisation of ctx 1) orders the
 * two streams on the device: its own stream first waits for the owner's pending work, and the owner's
 * stream then waits for the call's work, so the owner's later calls see the result without a host sync.
 */
#ifndef SSM_B200_H
#define SSM_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SSM_OK            0
#define SSM_E_ARG        -1   /* null pointer / negative size                                           */
#define SSM_E_SHAPE      -2   /* DimensionMismatch: operands of incompatible sizes                      */
#define SSM_E_UNSUPPORTED -3  /* (l,u,r) outside the compiled limits (SSM_MAX_*)                        */
#define SSM_E_STATE      -4   /* e.g. re-factorising an already factored object (semiseparableqr.jl:105) */
#define SSM_E_NOMEM      -5

#define SSM_MAX_BW   8        /* l, u (and BPS l, m) supported up to this by the generic kernels        */
#define SSM_MAX_RANK 4        /* r, p                                                                   */

#define SSM_DIST_DD 0         /* synthetic inputs, SURVEY 8(d): diagonally dominant band + small fill   */
#define SSM_DIST_BC 1         /* boundary-row (spectral) form: U=[I_r;0]                                */

typedef struct ssm_ctx   ssm_ctx;    /* device + stream + staging workspace                             */
typedef struct ssm_ab    ssm_ab;     /* batch of AlmostBandedMatrix operands, device resident           */
typedef struct ssm_abqr  ssm_abqr;   /* batch of QR{Float64,AlmostBandedMatrix} factorisations          */
typedef struct ssm_bps   ssm_bps;    /* batch of BandedPlusSemiseparableMatrix operands                 */
typedef struct ssm_bpsqr ssm_bpsqr;  /* batch of QR(BandedPlusSemiseparableMatrix, tau)                 */
typedef struct ssm_vec   ssm_vec;    /* batch of length-n vectors (right-hand sides / solutions)        */

/* ---- library / context ------------------------------------------------------------------------- */
int         ssm_version(void);
const char *ssm_last_error(void);
int  ssm_ctx_create(int device, ssm_ctx **ctx);                 /* own non-blocking stream               */
int  ssm_ctx_create_on_stream(int device, void *cuda_stream, ssm_ctx **ctx); /* caller's cudaStream_t   */
int  ssm_ctx_destroy(ssm_ctx *ctx);
int  ssm_ctx_sync(ssm_ctx *ctx);
/* tuning knobs (kernel variant selection for A/B measurements): key in {"ab_variant" (0 auto, 1 direct loads, 2 per-step
   bulk-async ring, 4 grouped ring with a producer warp, 9 generic shapes), "pipeline_stages", "group_bytes",
   "group_max_per_sm" (auto: batches of at most this many 32-system tiles per SM use variant 4), "group_smem_kb" (shared memory a
   variant-4 CTA may take; lower it, e.g. 96, when several contexts work on one GPU at the same time), and for the single-system
   chunked path: "abc_tiled_min_chunks" (chunk count from which the R records are tiled, default 12288), "abc_gs" (lanes per chunk of the
   lanes = columns stage 1 of the wide shapes, 4 | 8), "abc_lpc" (one-lane-per-chunk stage 1 of the shapes l+u = 8 | 6 | 4, r = 2:
   0 never, 1 from the second solve of the same operands on — it builds an interleaved copy of them —, 2 from the first)} */
int  ssm_ctx_set_option(ssm_ctx *ctx, const char *key, int value);
int  ssm_ctx_get_option(ssm_ctx *ctx, const char *key, int *value);
/* number of kernels launched by this ctx since creation (bench.py's gpu_launches) */
int64_t ssm_ctx_launch_count(ssm_ctx *ctx);
/* the cudaStream_t this context enqueues on, for callers that order their own device work against it (calls that take DEVICE pointers only
   enqueue: the caller's producer kernels must be ordered before, its consumers after — the Python twin does this for torch tensors) */
void *ssm_ctx_stream(ssm_ctx *ctx);
/* Device blocks of destroyed objects (<= 64 MiB each, <= 512 MiB and 256 blocks per device) are kept on a per-device free list and handed to the
   next ssm_*_create of a similar size, so that a caller who builds, factorises and destroys one matrix per call — what Julia's
   qr(A) / A \ b do (AlmostBandedMatrix.jl:126-133 allocates its factors per call as well) — pays no cudaMalloc / cudaFree.
   ssm_pool_trim returns the listed blocks of `device` to the driver (bytes freed); ssm_pool_stats reports the list
   (any output pointer may be null).  SSM_POOL=0 in the environment disables the list. */
int64_t ssm_pool_trim(int device);
int  ssm_pool_stats(int device, int64_t *bytes, int64_t *blocks, int64_t *hits, int64_t *misses);

/* ---- vectors ------------------------------------------------------------------------------------ */
int  ssm_vec_create(ssm_ctx *ctx, int n, int64_t nb, ssm_vec **v);
int  ssm_vec_destroy(ssm_vec *v);
int  ssm_vec_set(ssm_vec *v, const double *b, int64_t stride);  /* reference layout -> device layout     */
int  ssm_vec_get(const ssm_vec *v, double *b, int64_t stride);  /* device layout -> reference layout     */
int  ssm_vec_copy(ssm_vec *dst, const ssm_vec *src);
int  ssm_vec_generate(ssm_vec *v, uint64_t seed, int array_id); /* b = 2u-1 of the shared generator      */

/* ---- AlmostBandedMatrix (src/AlmostBandedMatrix.jl) ------------------------------------------------ */
/* AlmostBandedMatrix(bands, fill) constructor + size check (:7-18): allocates the device batch          */
int  ssm_ab_create(ssm_ctx *ctx, int n, int l, int u, int r, int64_t nb, ssm_ab **A);
int  ssm_ab_destroy(ssm_ab *A);
int  ssm_ab_set(ssm_ab *A, const double *bands, int64_t bands_stride, const double *U, int64_t U_stride,
                const double *V, int64_t V_stride);
int  ssm_ab_get(const ssm_ab *A, double *bands, int64_t bands_stride, double *U, int64_t U_stride,
                double *V, int64_t V_stride);
int  ssm_ab_generate(ssm_ab *A, uint64_t seed, int dist);       /* device-side synthetic operands        */
/* resize(A, n_new, n_new) (AlmostBandedMatrix.jl:338-348, the step resizedata! :368-371 takes when a cached operand outgrows its
 * storage): a new batch of n_new x n_new matrices holding the same bands and fill; the new band entries, the new rows of U and the
 * new columns of V are zero.  n_new < n is a DimensionMismatch (SSM_E_SHAPE), as Zeros(m - size(U,1), r) is in the reference. */
int  ssm_ab_resize(const ssm_ab *A, int n_new, ssm_ab **out);
/* The copy resizedata! makes after growing (:373-389): rows [k0, k1) (0-based, half open) of every matrix — the band entries
 * A[i, i-l .. i+u], U[i, :] and V[:, i] — are taken from full reference-layout arrays of A's size (same arguments as ssm_ab_set);
 * all other rows keep their contents.  Host arrays: only the windows that hold those rows are copied to the device.          */
int  ssm_ab_set_rows(ssm_ab *A, int k0, int k1, const double *bands, int64_t bands_stride, const double *U, int64_t U_stride,
                     const double *V, int64_t V_stride);
/* Vcat(a, Bd) front end (AlmostBandedMatrix.jl:298-305 bandwidth rule, :316-337 copyto!): the boundary-row form spectral-method
 * callers produce.  a: r x n dense rows (column-major), Bd: (lb+ub+1) x n band storage of the (n-r) x n banded block.  Builds
 * U = [I_r; 0], V = a and the bands directly in the device layout.  A must have been created with r rows of fill and bandwidths
 * >= ssm_ab_vcat_bandwidths(r, lb, ub).                                                                                      */
int  ssm_ab_vcat_bandwidths(int r, int lb, int ub, int *l, int *u);
int  ssm_ab_set_vcat(ssm_ab *A, const double *a, int64_t a_stride, const double *Bd, int64_t b_stride, int lb, int ub);

/* qr(A) / factorize(A): _qr -> almostbanded_qr -> widen (:30-38) + _almostbanded_qr! (:125-179).
 * A is not mutated (test_almostbanded.jl:146).  *F must be NULL (a new factor object is returned) or an
 * existing factor object of the same shape, which is overwritten (steady-state refactorisation).        */
int  ssm_ab_qr(const ssm_ab *A, ssm_abqr **F);
/* _almostbanded_qr!(A, tau, ncols) (:139-172): form only the first ncols reflectors (0 <= ncols <= n).  Rows and columns
 * past ncols of F.factors hold the partially updated trailing matrix (widened band + updated fill U), tau[ncols+1:] = 0 —
 * what adaptive / incremental callers continue from.  ssm_ab_qr(A, F) == ssm_ab_qr_cols(A, n - 1, F) (:137).            */
int  ssm_ab_qr_cols(const ssm_ab *A, int ncols, ssm_abqr **F);
/* Continue a partial factorisation in place: F = ssm_ab_qr_cols(A, n0) holds n0 reflectors; afterwards it holds ncols >= n0 of them and equals
 * ssm_ab_qr_cols(A, ncols) (the same A) to rounding (bit for bit when both run the same kernel family).  The step adaptive QR repeats as it
 * needs more columns (:139-172 run on the trailing view).  ssm_abqr_ldiv / _solve / _upper_ldiv / _ldiv_mat need the complete factorisation
 * (ncols >= n - 1) and return SSM_E_STATE before that; ssm_abqr_lmul_q / _qt apply the reflectors formed so far.
 * A factor object shares V with the operand batch it was made from; writing to that batch afterwards (ssm_ab_set, _set_rows, _set_vcat,
 * _generate) first gives the batch a V of its own, so live factor objects keep the matrix they factored (qr(A) copies, :129-133). */
int  ssm_abqr_continue(ssm_abqr *F, const ssm_ab *A, int ncols);
/* Grow a partial factorisation together with its operand — the step adaptive QR takes when the cached operand has outgrown its storage
 * (resizedata!, :357-394) while a factorisation of its leading columns exists: F = ssm_ab_qr_cols(A_old, k0), A = the operand after
 * ssm_ab_resize + ssm_ab_set_rows (its old n0 x n0 block unchanged).  Afterwards F has A's size and equals ssm_ab_qr_cols(A, k0) to
 * rounding — the stored reflectors are applied to the new columns (materialize!(Lmul{AdjQRPackedQLayout}), :201-207, restricted to them) —
 * and ssm_abqr_continue(F, A, ncols) goes on from there.  k0 + l > n0 is SSM_E_STATE: such reflectors saw columns cut off by the old size. */
int  ssm_abqr_resize(ssm_abqr *F, const ssm_ab *A);
int  ssm_abqr_destroy(ssm_abqr *F);
/* F.factors (widened band storage (2l+u+1) x n), F.factors.fill U (n x r), F.tau (n) (:174-185)           */
int  ssm_abqr_get(const ssm_abqr *F, double *factors, int64_t f_stride, double *U, int64_t U_stride,
                  double *tau, int64_t tau_stride);
/* The inverse of ssm_abqr_get: a factor object made from the reference's own fields — F.factors (widened band storage (2l+u+1) x n, its
 * fill U n x r and V r x n) and F.tau (:174-185) — for a QR{Float64,AlmostBandedMatrix} this library instance did not produce (a copy, a
 * deserialised factorisation), so that ldiv!(F, b) works on any F.  ncols < 0 means the complete factorisation (n - 1 reflectors).        */
int  ssm_abqr_set(ssm_ctx *ctx, int n, int l, int u, int r, int64_t nb, const double *factors, int64_t f_stride, const double *U,
                  int64_t U_stride, const double *V, int64_t V_stride, const double *tau, int64_t tau_stride, int ncols, ssm_abqr **F);
/* ldiv!(F, b) (:187-192): b <- R \ (Q'b), in place                                                        */
int  ssm_abqr_ldiv(const ssm_abqr *F, ssm_vec *b);
/* F \ b = ldiv!(F, copy(b)): x <- R \ (Q'b), b untouched (x may alias b)                                   */
int  ssm_abqr_solve(const ssm_abqr *F, const ssm_vec *b, ssm_vec *x);
/* ldiv!(F, B::AbstractMatrix) (:194-199): several right-hand sides per system, one pass over the factors per group of 4 or 8
 * columns.  A matrix batch holds, per system, the Julia Matrix memory (n x nrhs column-major).                               */
typedef struct ssm_mat ssm_mat;
int  ssm_mat_create(ssm_ctx *ctx, int n, int nrhs, int64_t nb, ssm_mat **B);
int  ssm_mat_destroy(ssm_mat *B);
int  ssm_mat_set(ssm_mat *B, const double *data, int64_t stride);
int  ssm_mat_get(const ssm_mat *B, double *data, int64_t stride);
int  ssm_abqr_ldiv_mat(const ssm_abqr *F, ssm_mat *B);
/* lmul!(F.Q', b) / lmul!(F.Q, b), Q = QRPackedQ(bandpart(F.factors), F.tau) (:181,201-207)                */
int  ssm_abqr_lmul_qt(const ssm_abqr *F, ssm_vec *b);
int  ssm_abqr_lmul_q(const ssm_abqr *F, ssm_vec *b);
/* ldiv!(UpperTriangular(F.R), b): _almostbanded_upper_ldiv! (:215-248)                                    */
int  ssm_abqr_upper_ldiv(const ssm_abqr *F, ssm_vec *b);
/* A \ b (:125 _factorize = qr, then ldiv!): fused factor + solve, x <- A \ b (x may alias b); the factors
 * are not kept (R still round-trips HBM in a scratch buffer owned by A).                                 */
int  ssm_ab_solve(const ssm_ab *A, const ssm_vec *b, ssm_vec *x);
/* Triangular solves on an unfactored AlmostBandedMatrix (:242-270): upper!=0 -> (Unit)UpperTriangular(A)\b */
int  ssm_ab_tri_ldiv(const ssm_ab *A, ssm_vec *b, int upper, int unit);
/* relres[s] = ||b - A x||_2 / ||b||_2 via getindex semantics (:84-90), one value per system               */
int  ssm_ab_residual(const ssm_ab *A, const ssm_vec *x, const ssm_vec *b, double *relres);
/* One-shot HOST-buffer path (end-to-end): H2D + pack + qr + ldiv! + D2H, chunk-pipelined over streams.    */
int  ssm_ab_solve_host(ssm_ctx *ctx, int n, int l, int u, int r, int64_t nb, const double *bands,
                       const double *U, const double *V, const double *b, double *x);
/* Batch sharder (north_star: independent systems split across the GPUs of one box, no NCCL; the host buffers are the
 * gather).  Shard g of G owns systems [s0, s1), cut at multiples of 32; ssm_ab_solve_host_sharded runs ssm_ab_solve_host on
 * every listed device from its own host thread and returns when all shards are done.                                   */
int  ssm_shard_range(int64_t nb, int g, int G, int64_t *s0, int64_t *s1);
int  ssm_ab_solve_host_sharded(int ndev, const int *devices, int n, int l, int u, int r, int64_t nb, const double *bands,
                               const double *U, const double *V, const double *b, double *x);

/* ---- BandedPlusSemiseparableMatrix (src/BandedPlusSemiseparableMatrix.jl, src/semiseparableqr.jl) -- */
/* BandedPlusSemiseparableMatrix(B,(U,V),(W,S)) + size check (BandedPlusSemiseparableMatrix.jl:10-16)      */
int  ssm_bps_create(ssm_ctx *ctx, int n, int l, int m, int r, int p, int64_t nb, ssm_bps **A);
int  ssm_bps_destroy(ssm_bps *A);
int  ssm_bps_set(ssm_bps *A, const double *B, int64_t B_stride, const double *U, const double *V,
                 int64_t UV_stride, const double *W, const double *S, int64_t WS_stride);
int  ssm_bps_generate(ssm_bps *A, uint64_t seed);
/* qr(A) (semiseparableqr.jl:104-152): returns QR(BandedPlusSemiseparableMatrix(F.factors), tau)           */
int  ssm_bps_qr(const ssm_bps *A, ssm_bpsqr **F);
int  ssm_bpsqr_destroy(ssm_bpsqr *F);
/* F.factors = BPS(B (l, l+m), (U, Vnew), (W [alpha beta], S [S AtU])) and tau                              */
int  ssm_bpsqr_get(const ssm_bpsqr *F, double *B, int64_t B_stride, double *U, double *V, int64_t UV_stride,
                   double *W, double *S, int64_t WS_stride, double *tau, int64_t tau_stride);
/* lmul!(F.Q', b) (semiseparableqr.jl:863-905)                                                              */
int  ssm_bpsqr_lmul_qt(const ssm_bpsqr *F, ssm_vec *b);
/* ldiv!(UpperTriangular(F.factors), b) (BandedPlusSemiseparableMatrix.jl:53-70)                            */
int  ssm_bpsqr_upper_ldiv(const ssm_bpsqr *F, ssm_vec *b);
/* F \ b = ldiv!(F.R, lmul!(F.Q', b)) (the two halves test_qr.jl:47-53 exercises)                           */
int  ssm_bpsqr_ldiv(const ssm_bpsqr *F, ssm_vec *b);
int  ssm_bpsqr_solve(const ssm_bpsqr *F, const ssm_vec *b, ssm_vec *x);   /* x <- F \ b, b untouched (x may alias b) */
/* rq_mul(F.factors, F.tau) (semiseparableqr.jl:170-311, 642-845): R*Q = Q'AQ of a SYMMETRIC BandedPlusSemiseparableMatrix (W = V,
 * S = U, m = l), returned as a new symmetric batch BPS(B (l, l), (U_new, V_new), (V_new, U_new)) — one step of the batched QR
 * iteration.  *RQ must be NULL (a new batch is returned) or a previous result of the same shape, which is overwritten.
 * SSM_E_SHAPE when the factors are not those of a symmetric shape, SSM_E_STATE when S[:, 1:r] != U (:195-205).                     */
int  ssm_bpsqr_rq_mul(const ssm_bpsqr *F, ssm_bps **RQ);
/* the operands of a batch back in the reference layout (any pointer may be NULL)                                                 */
int  ssm_bps_get(const ssm_bps *A, double *B, int64_t B_stride, double *U, double *V, int64_t UV_stride, double *W, double *S,
                 int64_t WS_stride);
/* A * x (BandedPlusSemiseparableMatrix.jl:21-41) and the residual built on it                              */
int  ssm_bps_mul(const ssm_bps *A, const ssm_vec *x, ssm_vec *y);
int  ssm_bps_residual(const ssm_bps *A, const ssm_vec *x, const ssm_vec *b, double *relres);
/* ldiv!(UpperTriangular(A), b) on an unfactored BPS (test_bandedplussemiseparable.jl:17)                   */
int  ssm_bps_upper_ldiv(const ssm_bps *A, ssm_vec *b);

/* ---- one LARGE AlmostBandedMatrix system, chunked along the n axis (BASELINE config 4) ------------------------
 * `A \ b` (AlmostBandedMatrix.jl:125-133,187-199) for a single system whose sequential sweep cannot fill a GPU:
 * a partitioned orthogonal factorisation over P row chunks.  Returns the reference's SOLUTION
 * (rel. error <= 1e-12 on well-conditioned systems), not its factors.  Operands in the reference layout, host or device. */
typedef struct ssm_abc ssm_abc;
int  ssm_abc_create(ssm_ctx *ctx, int64_t n, int l, int u, int r, ssm_abc **A);
int  ssm_abc_destroy(ssm_abc *A);
int  ssm_abc_set(ssm_abc *A, const double *bands, const double *U, const double *V);
int  ssm_abc_generate(ssm_abc *A, uint64_t seed, int dist);            /* system index 0 of the shared generator  */
int  ssm_abc_vec_generate(ssm_abc *A, double *b_dev, uint64_t seed);   /* b = 2u-1 (array id 3), device pointer    */
/* x <- A \ b; b, x: n doubles, host or device, x may alias b; nchunks = 0 picks ~2048 rows per chunk            */
int  ssm_abc_solve(ssm_abc *A, const double *b, double *x, int nchunks);
int  ssm_abc_residual(ssm_abc *A, const double *x, const double *b, double *relres);
int  ssm_abc_chunks(const ssm_abc *A);                                 /* P of the last solve                      */
/* inspection for tests: what = 0 level-0 reduced blocks, 1 separator solution, 2 R-row records, 3 operand records */
int  ssm_abc_debug_get(ssm_abc *A, int what, double *out, int64_t count);

#ifdef __cplusplus
}
#endif
#endif /* SSM_B200_H */
