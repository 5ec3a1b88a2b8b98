// api.cu — the C ABI of libssmb200.so (include/ssm_b200.h): handle management, host/device pointer staging and
// the AlmostBanded entry points.  BPS entry points live in bps_api.cu.
#include <cstdarg>
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include "ssm_internal.cuh"

namespace ssm {

static thread_local std::string g_err;

void set_error(const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}
int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    cudaGetLastError();
    return (int)e;
}

// ---- per-device free list of small device blocks ---------------------------------------------------------------
// A caller that factorises one matrix per call (the Julia glue: qr(A) creates an operand batch, a factor object and a vector and
// destroys them again) paid a cudaMalloc + cudaFree per buffer, several times the kernels' own time at n ~ 1,000.  Blocks of at
// most POOL_BLOCK_MAX bytes therefore go back to a per-device list (at most POOL_TOTAL_MAX bytes and POOL_MAX_BLOCKS blocks, oldest evicted first) and
// dev_alloc takes the smallest listed block of at least — and at most twice — the requested size.  A block is only listed when
// the stream it was allocated for is idle (every ssm_*_destroy synchronises its context's stream first), so it can be handed to any
// other stream; otherwise it is cudaFree'd as before.  Large batches never pass through the list: their allocation cost is
// invisible beside their kernels and holding them back would only take memory from the next batch.
// SSM_POOL=0 in the environment switches the list off; ssm_pool_trim() gives the listed blocks back to the driver.
namespace {
constexpr size_t POOL_BLOCK_MAX = 64ull << 20;
constexpr size_t POOL_TOTAL_MAX = 512ull << 20;
constexpr int    POOL_MAX_DEV   = 64;
constexpr size_t POOL_MAX_BLOCKS = 256;          // the list is searched linearly
struct PoolEntry { double *p; size_t cap; };
struct DevPool { std::mutex mu; std::vector<PoolEntry> blocks; size_t bytes = 0; int64_t hits = 0, misses = 0; };
DevPool *pool_of(int device) {
    static DevPool *pools = new DevPool[POOL_MAX_DEV];      // never destroyed: no cudaFree after the runtime has shut down
    static const bool on = [] { const char *e = getenv("SSM_POOL"); return !(e && e[0] == '0'); }();
    return (on && device >= 0 && device < POOL_MAX_DEV) ? &pools[device] : nullptr;
}
void raw_free(int device, double *p) {
    int cur = 0; cudaGetDevice(&cur);
    if (cur != device) cudaSetDevice(device);
    cudaFree(p);
    if (cur != device) cudaSetDevice(cur);
}
double *pool_take(int device, size_t count, size_t *cap) {
    DevPool *P = pool_of(device);
    if (!P || count * sizeof(double) > POOL_BLOCK_MAX) return nullptr;
    std::lock_guard<std::mutex> g(P->mu);
    int best = -1;
    for (int i = 0; i < (int)P->blocks.size(); i++) {
        const size_t c = P->blocks[i].cap;
        if (c >= count && c <= 2 * count + 512 && (best < 0 || c < P->blocks[best].cap)) best = i;
    }
    if (best < 0) { P->misses++; return nullptr; }
    PoolEntry e = P->blocks[best];
    P->blocks.erase(P->blocks.begin() + best);
    P->bytes -= e.cap * sizeof(double);
    P->hits++;
    *cap = e.cap;
    return e.p;
}
bool pool_give(int device, double *p, size_t cap) {
    DevPool *P = pool_of(device);
    const size_t bytes = cap * sizeof(double);
    if (!P || bytes > POOL_BLOCK_MAX) return false;
    std::vector<PoolEntry> evict;
    {
        std::lock_guard<std::mutex> g(P->mu);
        while ((P->bytes + bytes > POOL_TOTAL_MAX || P->blocks.size() >= POOL_MAX_BLOCKS) && !P->blocks.empty()) {
            evict.push_back(P->blocks.front());
            P->bytes -= P->blocks.front().cap * sizeof(double);
            P->blocks.erase(P->blocks.begin());
        }
        P->blocks.push_back({p, cap});
        P->bytes += bytes;
    }
    for (auto &e : evict) raw_free(device, e.p);
    return true;
}
int64_t pool_trim_dev(int device) {
    DevPool *P = pool_of(device);
    if (!P) return 0;
    std::vector<PoolEntry> all;
    { std::lock_guard<std::mutex> g(P->mu); all.swap(P->blocks); P->bytes = 0; }
    int64_t freed = 0;
    for (auto &e : all) { raw_free(device, e.p); freed += (int64_t)(e.cap * sizeof(double)); }
    return freed;
}
}  // namespace

DevBuf::~DevBuf() {
    if (!p) return;
    // listed only when nothing can still be using it (see above); cudaFree waits for the device, the list does not.
    // A retired context synchronised its stream when it was destroyed and nothing can have been enqueued since.
    const bool idle = !foreign.load() && (!tok || !tok->alive.load() || cudaStreamQuery(tok->st) == cudaSuccess);
    if (idle && pool_give(device, p, cap)) return;
    cudaGetLastError();
    raw_free(device, p);
}
void ctx_retain(ssm_ctx *c) { c->live.fetch_add(1); }
void ctx_release(ssm_ctx *c) {
    if (c->live.fetch_sub(1) == 1 && c->retired.load()) delete c;
}
int dev_alloc(ssm_ctx *c, size_t count, bool zero, DevBufPtr &out) {
    const int device = c->device;
    cudaStream_t st = c->stream;
    auto b = std::make_shared<DevBuf>();
    b->device = device; b->count = count; b->tok = c->tok;
    if (count == 0) count = 1;
    b->cap = count;
    b->p = pool_take(device, count, &b->cap);
    if (!b->p) {
        cudaError_t e = cudaMalloc(&b->p, count * sizeof(double));
        if (e != cudaSuccess && pool_trim_dev(device) > 0) { cudaGetLastError(); e = cudaMalloc(&b->p, count * sizeof(double)); }
        if (e != cudaSuccess) { b->p = nullptr; set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(double), cudaGetErrorString(e)); cudaGetLastError(); return SSM_E_NOMEM; }
    }
    if (zero) SSM_CUDA(cudaMemsetAsync(b->p, 0, count * sizeof(double), st));
    out = b;
    return 0;
}

bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int staging_dev(ssm_ctx *c, size_t count, double **out) {
    if (c->stg.dev_count < count) {
        if (c->stg.dev) { SSM_CUDA(cudaStreamSynchronize(c->stream)); SSM_CUDA(cudaFree(c->stg.dev)); c->stg.dev = nullptr; c->stg.dev_count = 0; }
        cudaError_t e = cudaMalloc(&c->stg.dev, count * sizeof(double));
        if (e != cudaSuccess) { set_error("staging cudaMalloc of %zu bytes failed", count * sizeof(double)); cudaGetLastError(); return SSM_E_NOMEM; }
        c->stg.dev_count = count;
    }
    *out = c->stg.dev;
    return 0;
}

// copy `count` systems of `elems` doubles each, source stride `stride`, into a dense device block
static int h2d_block(ssm_ctx *c, double *dst, const double *src, int64_t stride, size_t elems, int64_t count) {
    if (elems == 0 || count == 0) return 0;
    if ((size_t)stride == elems) SSM_CUDA(cudaMemcpyAsync(dst, src, elems * count * sizeof(double), cudaMemcpyDefault, c->stream));
    else SSM_CUDA(cudaMemcpy2DAsync(dst, elems * sizeof(double), src, stride * sizeof(double), elems * sizeof(double), count, cudaMemcpyDefault, c->stream));
    return 0;
}
static int d2h_block(ssm_ctx *c, double *dst, int64_t stride, const double *src, size_t elems, int64_t count) {
    if (elems == 0 || count == 0) return 0;
    if ((size_t)stride == elems) SSM_CUDA(cudaMemcpyAsync(dst, src, elems * count * sizeof(double), cudaMemcpyDefault, c->stream));
    else SSM_CUDA(cudaMemcpy2DAsync(dst, stride * sizeof(double), src, elems * sizeof(double), elems * sizeof(double), count, cudaMemcpyDefault, c->stream));
    return 0;
}

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
    ~DeviceGuard() { int cur; cudaGetDevice(&cur); if (cur != prev) cudaSetDevice(prev); }
};

static const int64_t STAGE_BYTES_MAX = 256ll << 20;
static int64_t chunk_systems(size_t doubles_per_system, int64_t nb) {
    int64_t ch = STAGE_BYTES_MAX / (int64_t)(doubles_per_system * sizeof(double) + 1);
    ch = std::max<int64_t>(TILE, (ch / TILE) * TILE);
    return std::min<int64_t>(ch, ((nb + TILE - 1) / TILE) * TILE);
}

}  // namespace ssm

using namespace ssm;

#define SSM_CHECK_ARG(cond, msg) do { if (!(cond)) { set_error("%s: %s", __func__, msg); return SSM_E_ARG; } } while (0)

extern "C" {

int ssm_version(void) { return 100; }
const char *ssm_last_error(void) { return g_err.c_str(); }

static int ctx_common(int device, ssm_ctx *c) {
    int ndev = 0;
    SSM_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("device %d out of range (%d visible)", device, ndev); return SSM_E_ARG; }
    c->device = device;
    SSM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SSM_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) { set_error("libssmb200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor); return SSM_E_UNSUPPORTED; }
    return 0;
}
int ssm_ctx_create(int device, ssm_ctx **out) {
    SSM_CHECK_ARG(out, "null output");
    auto *c = new ssm_ctx();
    int rc = ctx_common(device, c);
    if (rc) { delete c; return rc; }
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
    c->own_stream = true;
    c->tok = std::make_shared<StreamToken>(); c->tok->st = c->stream;
    *out = c;
    return 0;
}
int ssm_ctx_create_on_stream(int device, void *cuda_stream, ssm_ctx **out) {
    SSM_CHECK_ARG(out, "null output");
    auto *c = new ssm_ctx();
    int rc = ctx_common(device, c);
    if (rc) { delete c; return rc; }
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    c->tok = std::make_shared<StreamToken>(); c->tok->st = c->stream;
    *out = c;
    return 0;
}
int ssm_ctx_destroy(ssm_ctx *c) {
    if (!c || c->retired.load()) return 0;
    DeviceGuard g(c->device);
    cudaStreamSynchronize(c->stream);
    hostpipe_free(c);
    abc_cache_free(c);
    if (c->stg.dev) { cudaFree(c->stg.dev); c->stg.dev = nullptr; c->stg.dev_count = 0; }
    if (c->stg.pin) { cudaFreeHost(c->stg.pin); c->stg.pin = nullptr; c->stg.pin_count = 0; }
    for (auto &s : c->aux) if (s) { cudaStreamDestroy(s); s = nullptr; }
    if (c->ev) { cudaEventDestroy(c->ev); c->ev = nullptr; }
    if (c->tok) c->tok->alive.store(false);          // blocks that outlive the context must not query its stream any more
    if (c->own_stream) cudaStreamDestroy(c->stream);
    // objects still alive (destroyed out of order by the host runtime): the struct stays until the last of them is gone
    ctx_retain(c);
    c->retired.store(true);
    ctx_release(c);
    return 0;
}
int ssm_ctx_sync(ssm_ctx *c) {
    SSM_CHECK_ARG(c, "null ctx");
    DeviceGuard g(c->device);
    SSM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int ssm_ctx_set_option(ssm_ctx *c, const char *key, int value) {
    SSM_CHECK_ARG(c && key, "null argument");
    if (!strcmp(key, "ab_variant")) c->ab_variant = value;
    else if (!strcmp(key, "pipeline_stages")) c->pipeline_stages = value;
    else if (!strcmp(key, "qt_stages")) c->qt_stages = value;
    else if (!strcmp(key, "bs_stages")) c->bs_stages = value;
    else if (!strcmp(key, "group_max_per_sm")) c->group_max_per_sm = value;
    else if (!strcmp(key, "group_bytes")) c->group_bytes = value;
    else if (!strcmp(key, "group_smem_kb")) c->group_smem_kb = value;
    else if (!strcmp(key, "carveout_all")) c->carveout_all = value;
    else if (!strcmp(key, "abc_tiled_min_chunks")) c->abc_tiled_min_chunks = value;
    else if (!strcmp(key, "abc_lpc")) c->abc_lpc = value;
    else if (!strcmp(key, "abc_gs")) { SSM_CHECK_ARG(value == 4 || value == 8, "abc_gs must be 4 or 8"); c->abc_gs = value; }
    else { set_error("unknown option '%s'", key); return SSM_E_ARG; }
    return 0;
}
int ssm_ctx_get_option(ssm_ctx *c, const char *key, int *value) {
    SSM_CHECK_ARG(c && key && value, "null argument");
    if (!strcmp(key, "ab_variant")) *value = c->ab_variant;
    else if (!strcmp(key, "pipeline_stages")) *value = c->pipeline_stages;
    else if (!strcmp(key, "group_max_per_sm")) *value = c->group_max_per_sm;
    else if (!strcmp(key, "group_bytes")) *value = c->group_bytes;
    else if (!strcmp(key, "group_smem_kb")) *value = c->group_smem_kb;
    else if (!strcmp(key, "abc_tiled_min_chunks")) *value = c->abc_tiled_min_chunks;
    else if (!strcmp(key, "abc_gs")) *value = c->abc_gs;
    else if (!strcmp(key, "abc_lpc")) *value = c->abc_lpc;
    else if (!strcmp(key, "sm_count")) *value = c->sm_count;
    else { set_error("unknown option '%s'", key); return SSM_E_ARG; }
    return 0;
}
int64_t ssm_ctx_launch_count(ssm_ctx *c) { return c ? c->launches : 0; }
void *ssm_ctx_stream(ssm_ctx *c) { return c ? (void *)c->stream : nullptr; }
int64_t ssm_pool_trim(int device) { return pool_trim_dev(device); }
int ssm_pool_stats(int device, int64_t *bytes, int64_t *blocks, int64_t *hits, int64_t *misses) {
    DevPool *P = pool_of(device);
    int64_t v[4] = {0, 0, 0, 0};
    if (P) { std::lock_guard<std::mutex> g(P->mu); v[0] = (int64_t)P->bytes; v[1] = (int64_t)P->blocks.size(); v[2] = P->hits; v[3] = P->misses; }
    if (bytes) *bytes = v[0];
    if (blocks) *blocks = v[1];
    if (hits) *hits = v[2];
    if (misses) *misses = v[3];
    return 0;                         // list switched off (SSM_POOL=0): all zero
}


// ---- vectors ----------------------------------------------------------------------------------------------
int ssm_vec_create(ssm_ctx *c, int n, int64_t nb, ssm_vec **out) {
    SSM_CHECK_ARG(c && out, "null argument");
    SSM_CHECK_ARG(n >= 1 && nb >= 1, "n and nb must be positive");
    DeviceGuard g(c->device);
    auto *v = new ssm_vec();
    v->ctx = c; v->ref.bind(c); v->n = n; v->nb = nb; v->ntiles = ntiles_of(nb); v->np = (int64_t)n + PAD;
    int rc = dev_alloc(c, (size_t)(v->ntiles * v->np * 32), true, v->d);
    if (rc) { delete v; return rc; }
    *out = v;
    return 0;
}
int ssm_vec_destroy(ssm_vec *v) { if (v) { obj_sync(v->ctx); delete v; } return 0; }

int ssm_vec_set(ssm_vec *v, const double *b, int64_t stride) {
    SSM_CHECK_ARG(v && b, "null argument");
    ssm_ctx *c = v->ctx;
    DeviceGuard g(c->device);
    if (stride == 0) stride = v->n;
    if (stride < v->n) { set_error("ssm_vec_set: stride %lld < n", (long long)stride); return SSM_E_SHAPE; }
    if (is_device_ptr(b)) return launch_vec_pack(c, v, b, stride, 0, v->nb);
    const int64_t ch = chunk_systems((size_t)v->n, v->nb);
    for (int64_t s0 = 0; s0 < v->nb; s0 += ch) {
        const int64_t cnt = std::min(ch, v->nb - s0);
        double *stg;
        SSM_TRY(staging_dev(c, (size_t)cnt * v->n, &stg));
        SSM_TRY(h2d_block(c, stg, b + s0 * stride, stride, (size_t)v->n, cnt));
        SSM_TRY(launch_vec_pack(c, v, stg, v->n, s0, cnt));
    }
    SSM_CUDA(cudaStreamSynchronize(c->stream));   // caller's host buffer may be reused on return
    return 0;
}
int ssm_vec_get(const ssm_vec *v, double *b, int64_t stride) {
    SSM_CHECK_ARG(v && b, "null argument");
    ssm_ctx *c = v->ctx;
    DeviceGuard g(c->device);
    if (stride == 0) stride = v->n;
    if (stride < v->n) { set_error("ssm_vec_get: stride %lld < n", (long long)stride); return SSM_E_SHAPE; }
    if (is_device_ptr(b)) return launch_vec_unpack(c, v, b, stride, 0, v->nb);
    const int64_t ch = chunk_systems((size_t)v->n, v->nb);
    for (int64_t s0 = 0; s0 < v->nb; s0 += ch) {
        const int64_t cnt = std::min(ch, v->nb - s0);
        double *stg;
        SSM_TRY(staging_dev(c, (size_t)cnt * v->n, &stg));
        SSM_TRY(launch_vec_unpack(c, v, stg, v->n, s0, cnt));
        SSM_TRY(d2h_block(c, b + s0 * stride, stride, stg, (size_t)v->n, cnt));
    }
    SSM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int ssm_vec_copy(ssm_vec *dst, const ssm_vec *src) {
    SSM_CHECK_ARG(dst && src, "null argument");
    if (dst->n != src->n || dst->nb != src->nb) { set_error("ssm_vec_copy: shape mismatch"); return SSM_E_SHAPE; }
    DeviceGuard g(dst->ctx->device);
    SSM_CUDA(cudaMemcpyAsync(dst->d->p, src->d->p, src->d->count * sizeof(double), cudaMemcpyDeviceToDevice, dst->ctx->stream));
    // a copy between contexts reads src on dst's stream: finish it here, so that src's own stream still orders everything that touches src
    if (dst->ctx->stream != src->ctx->stream) SSM_CUDA(cudaStreamSynchronize(dst->ctx->stream));
    return 0;
}
int ssm_vec_generate(ssm_vec *v, uint64_t seed, int array_id) {
    SSM_CHECK_ARG(v, "null argument");
    DeviceGuard g(v->ctx->device);
    return launch_vec_generate(v->ctx, v, seed, array_id);
}

// ---- AlmostBanded -------------------------------------------------------------------------------------------
int ssm_ab_create(ssm_ctx *c, int n, int l, int u, int r, int64_t nb, ssm_ab **out) {
    SSM_CHECK_ARG(c && out, "null argument");
    SSM_CHECK_ARG(n >= 1 && nb >= 1, "n and nb must be positive");
    SSM_CHECK_ARG(l >= 0 && u >= 0 && r >= 0, "bandwidths and rank must be non-negative");
    if (l > SSM_MAX_BW || u > SSM_MAX_BW || r > SSM_MAX_RANK) { set_error("(l,u,r)=(%d,%d,%d) exceeds compiled limits (%d,%d,%d)", l, u, r, SSM_MAX_BW, SSM_MAX_BW, SSM_MAX_RANK); return SSM_E_UNSUPPORTED; }
    DeviceGuard g(c->device);
    auto *A = new ssm_ab();
    A->ctx = c; A->ref.bind(c); A->n = n; A->l = l; A->u = u; A->r = r; A->nb = nb; A->ntiles = ntiles_of(nb); A->np = (int64_t)n + PAD;
    int rc = dev_alloc(c, (size_t)(A->ntiles * A->np * (l + u + 1 + r) * 32), true, A->AU);
    if (!rc) rc = dev_alloc(c, (size_t)(A->ntiles * A->np * (r > 0 ? r : 1) * 32), true, A->V);
    if (rc) { delete A; return rc; }
    *out = A;
    return 0;
}
int ssm_ab_destroy(ssm_ab *A) { if (A) { obj_sync(A->ctx); delete A; } return 0; }

// resize(A, n_new, n_new) (AlmostBandedMatrix.jl:338-348): the same bands and fill in a larger matrix, new band rows / columns, new rows of U
// and new columns of V zero.  A tile's rows are contiguous ([tile][np][nf][32]), so the old rows of every tile move in one strided copy into
// the zero-initialised larger batch; nothing is repacked.
static int copy_tile_rows(ssm_ctx *c, double *dst, int64_t np_dst, const double *src, int64_t np_src, int64_t rows, int nf, int64_t ntiles) {
    const size_t rec = (size_t)nf * TILE * sizeof(double);
    if ((size_t)np_dst * rec < (size_t)0x7fffffff)
        SSM_CUDA(cudaMemcpy2DAsync(dst, (size_t)np_dst * rec, src, (size_t)np_src * rec, (size_t)rows * rec, (size_t)ntiles, cudaMemcpyDeviceToDevice, c->stream));
    else
        for (int64_t t = 0; t < ntiles; t++)
            SSM_CUDA(cudaMemcpyAsync(dst + t * np_dst * nf * TILE, src + t * np_src * nf * TILE, (size_t)rows * rec, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}
int ssm_ab_resize(const ssm_ab *A, int n_new, ssm_ab **out) {
    SSM_CHECK_ARG(A && out, "null argument");
    if (n_new < A->n) { set_error("ssm_ab_resize: DimensionMismatch: cannot shrink a %d x %d matrix to %d x %d", A->n, A->n, n_new, n_new); return SSM_E_SHAPE; }
    ssm_ctx *c = A->ctx;
    ssm_ab *B = nullptr;
    SSM_TRY(ssm_ab_create(c, n_new, A->l, A->u, A->r, A->nb, &B));
    DeviceGuard g(c->device);
    int rc = copy_tile_rows(c, B->AU->p, B->np, A->AU->p, A->np, A->n, A->l + A->u + 1 + A->r, A->ntiles);
    if (!rc && A->r > 0) rc = copy_tile_rows(c, B->V->p, B->np, A->V->p, A->np, A->n, A->r, A->ntiles);
    if (rc) { ssm_ab_destroy(B); return rc; }
    *out = B;
    return 0;
}

// A factor object made from A shares A's V block (it is never written by the path).  Before anything rewrites A's V the batch takes a block of
// its own — with the old contents when only some rows are about to change — so that live factor objects keep the matrix they factored.
static int ab_own_v(ssm_ab *A, bool keep) {
    if (!A->V || A->V.use_count() <= 1) return 0;
    ssm_ctx *c = A->ctx;
    DevBufPtr nv;
    SSM_TRY(dev_alloc(c, A->V->count, !keep, nv));
    if (keep) SSM_CUDA(cudaMemcpyAsync(nv->p, A->V->p, A->V->count * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    uses(A->V, c);
    A->V = nv;
    return 0;
}

int ssm_ab_set(ssm_ab *A, const double *bands, int64_t sb, const double *U, int64_t su, const double *V, int64_t sv) {
    SSM_CHECK_ARG(A && bands && (A->r == 0 || (U && V)), "null argument");
    ssm_ctx *c = A->ctx;
    DeviceGuard g(c->device);
    const size_t eb = (size_t)(A->l + A->u + 1) * A->n, eu = (size_t)A->n * A->r;
    if (sb == 0) sb = (int64_t)eb;
    if (su == 0) su = (int64_t)eu;
    if (sv == 0) sv = (int64_t)eu;
    if ((size_t)sb < eb || (size_t)su < eu || (size_t)sv < eu) { set_error("ssm_ab_set: Data and fill must be compatible size (stride smaller than one system)"); return SSM_E_SHAPE; }
    SSM_TRY(ab_own_v(A, false));
    const bool db = is_device_ptr(bands), du = A->r == 0 || is_device_ptr(U), dv = A->r == 0 || is_device_ptr(V);
    if (db && du && dv) return launch_ab_pack(c, A, bands, sb, U, su, V, sv, 0, A->nb);
    const int64_t ch = chunk_systems(eb + 2 * eu, A->nb);
    for (int64_t s0 = 0; s0 < A->nb; s0 += ch) {
        const int64_t cnt = std::min(ch, A->nb - s0);
        double *stg;
        SSM_TRY(staging_dev(c, (size_t)cnt * (eb + 2 * eu), &stg));
        double *sB = stg, *sU = sB + (size_t)cnt * eb, *sV = sU + (size_t)cnt * eu;
        SSM_TRY(h2d_block(c, sB, bands + s0 * sb, sb, eb, cnt));
        if (A->r > 0) { SSM_TRY(h2d_block(c, sU, U + s0 * su, su, eu, cnt)); SSM_TRY(h2d_block(c, sV, V + s0 * sv, sv, eu, cnt)); }
        SSM_TRY(launch_ab_pack(c, A, sB, (int64_t)eb, sU, (int64_t)eu, sV, (int64_t)eu, s0, cnt));
    }
    SSM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// Rows [k0, k1) of every matrix from full reference-layout arrays: the band entries A[i, i-l .. i+u], U[i, :] and V[:, i] for k0 <= i < k1;
// every other row keeps what it holds.  This is the copy resizedata! makes after it has grown the storage (AlmostBandedMatrix.jl:373-389:
// the new columns of V, the new rows of U, the new part of the band).  Host arrays: only the band columns [k0-l, k1+u), the rows
// k0..k1-1 of U and the columns k0..k1-1 of V cross the bus.
int ssm_ab_set_rows(ssm_ab *A, int k0, int k1, const double *bands, int64_t sb, const double *U, int64_t su, const double *V, int64_t sv) {
    SSM_CHECK_ARG(A && bands && (A->r == 0 || (U && V)), "null argument");
    if (k0 < 0 || k1 > A->n || k0 > k1) { set_error("ssm_ab_set_rows: BoundsError: rows [%d, %d) of a %d x %d matrix", k0, k1, A->n, A->n); return SSM_E_SHAPE; }
    if (k0 == k1) return 0;
    ssm_ctx *c = A->ctx;
    DeviceGuard g(c->device);
    const int n = A->n, l = A->l, u = A->u, r = A->r, nd = l + u + 1;
    const size_t eb = (size_t)nd * n, eu = (size_t)n * r;
    if (sb == 0) sb = (int64_t)eb;
    if (su == 0) su = (int64_t)eu;
    if (sv == 0) sv = (int64_t)eu;
    if ((size_t)sb < eb || (size_t)su < eu || (size_t)sv < eu) { set_error("ssm_ab_set_rows: Data and fill must be compatible size (stride smaller than one system)"); return SSM_E_SHAPE; }
    SSM_TRY(ab_own_v(A, true));
    const bool db = is_device_ptr(bands), du = r == 0 || is_device_ptr(U), dv = r == 0 || is_device_ptr(V);
    if (db && du && dv) return launch_ab_pack_rows(c, A, bands, sb, 0, U, su, n, V, sv, 0, 0, A->nb, k0, k1);
    const int64_t jlo = std::max(k0 - l, 0), jhi = std::min(k1 + u, n), rows = k1 - k0;
    const size_t wb = (size_t)nd * (jhi - jlo), wu = (size_t)rows * r;            // window sizes per system
    const int64_t ch = chunk_systems(wb + 2 * wu, A->nb);
    for (int64_t s0 = 0; s0 < A->nb; s0 += ch) {
        const int64_t cnt = std::min(ch, A->nb - s0);
        double *stg;
        SSM_TRY(staging_dev(c, (size_t)cnt * (wb + 2 * wu), &stg));
        double *sB = stg, *sU = sB + (size_t)cnt * wb, *sV = sU + (size_t)cnt * wu;
        SSM_TRY(h2d_block(c, sB, bands + s0 * sb + (int64_t)nd * jlo, sb, wb, cnt));
        for (int q = 0; q < r; q++)                                                // rows k0..k1-1 of column q of U: [system][q][row] in the window
            SSM_CUDA(cudaMemcpy2DAsync(sU + (size_t)q * rows, wu * sizeof(double), U + s0 * su + (int64_t)q * n + k0, (size_t)su * sizeof(double),
                                       (size_t)rows * sizeof(double), (size_t)cnt, cudaMemcpyDefault, c->stream));
        if (r > 0) SSM_TRY(h2d_block(c, sV, V + s0 * sv + (int64_t)r * k0, sv, wu, cnt));
        SSM_TRY(launch_ab_pack_rows(c, A, sB, (int64_t)wb, jlo, sU, (int64_t)wu, rows, sV, (int64_t)wu, k0, s0, cnt, k0, k1));
    }
    SSM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// almostbandwidths(Vcat(a, Bd)) (AlmostBandedMatrix.jl:298-303): (max(lb + r, r - 1), ub - r); the device batch needs u >= 0, a
// negative upper bandwidth is widened to 0 (the diagonal then holds the fill values it would otherwise have been read from).
int ssm_ab_vcat_bandwidths(int r, int lb, int ub, int *l, int *u) {
    SSM_CHECK_ARG(l && u && r >= 0, "bad arguments");
    *l = std::max(lb + r, r - 1);
    *u = std::max(ub - r, 0);
    return 0;
}

// AlmostBandedMatrix(Vcat(a, Bd)) = copyto!(AlmostBandedMatrix(undef, ...), Vcat(a, Bd)) (AlmostBandedMatrix.jl:48-49, 316-337):
// a = r x n dense rows (column-major), Bd = band storage of the (n-r) x n banded block with bandwidths (lb, ub).
int ssm_ab_set_vcat(ssm_ab *A, const double *a, int64_t a_stride, const double *Bd, int64_t b_stride, int lb, int ub) {
    SSM_CHECK_ARG(A && a && Bd, "null argument");
    ssm_ctx *c = A->ctx;
    DeviceGuard g(c->device);
    const int r = A->r;
    if (r < 1 || A->n <= r) { set_error("ssm_ab_set_vcat: need 1 <= r < n"); return SSM_E_SHAPE; }
    if (lb + ub + 1 < 1 || A->l < std::max(lb + r, r - 1) || A->u < std::max(ub - r, 0)) {
        set_error("ssm_ab_set_vcat: DimensionMismatch: bandwidths (%d,%d) of the batch cannot hold Vcat(%d rows, banded (%d,%d)); need (%d,%d)", A->l, A->u, r, lb, ub,
                  std::max(lb + r, r - 1), std::max(ub - r, 0));
        return SSM_E_SHAPE;
    }
    const size_t ea = (size_t)r * A->n, eb = (size_t)(lb + ub + 1) * A->n;
    if (a_stride == 0) a_stride = (int64_t)ea;
    if (b_stride == 0) b_stride = (int64_t)eb;
    if ((size_t)a_stride < ea || (size_t)b_stride < eb) { set_error("ssm_ab_set_vcat: stride smaller than one system"); return SSM_E_SHAPE; }
    SSM_TRY(ab_own_v(A, false));
    if (is_device_ptr(a) && is_device_ptr(Bd)) return launch_ab_pack_vcat(c, A, a, a_stride, Bd, b_stride, lb, ub);
    double *stg;
    SSM_TRY(staging_dev(c, (size_t)A->nb * (ea + eb), &stg));
    double *sA = stg, *sB = stg + (size_t)A->nb * ea;
    SSM_TRY(h2d_block(c, sA, a, a_stride, ea, A->nb));
    SSM_TRY(h2d_block(c, sB, Bd, b_stride, eb, A->nb));
    SSM_TRY(launch_ab_pack_vcat(c, A, sA, (int64_t)ea, sB, (int64_t)eb, lb, ub));
    SSM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// unpack helper shared by ssm_ab_get / ssm_abqr_get: run `launch` into device memory (direct or staged)
int ssm_ab_get(const ssm_ab *A, double *bands, int64_t sb, double *U, int64_t su, double *V, int64_t sv) {
    SSM_CHECK_ARG(A, "null argument");
    ssm_ctx *c = A->ctx;
    DeviceGuard g(c->device);
    const size_t eb = (size_t)(A->l + A->u + 1) * A->n, eu = (size_t)A->n * A->r;
    if (sb == 0) sb = (int64_t)eb;
    if (su == 0) su = (int64_t)eu;
    if (sv == 0) sv = (int64_t)eu;
    const bool anyhost = (bands && !is_device_ptr(bands)) || (U && !is_device_ptr(U)) || (V && !is_device_ptr(V));
    if (!anyhost) return launch_ab_unpack(c, A, bands, sb, U, su, V, sv);
    // host destination: unpack the whole batch into device staging, then copy out (test / inspection path)
    double *stg;
    SSM_TRY(staging_dev(c, (size_t)A->nb * (eb + 2 * eu), &stg));
    double *sB = stg, *sU = sB + (size_t)A->nb * eb, *sV = sU + (size_t)A->nb * eu;
    SSM_TRY(launch_ab_unpack(c, A, bands ? sB : nullptr, (int64_t)eb, U ? sU : nullptr, (int64_t)eu, V ? sV : nullptr, (int64_t)eu));
    if (bands) SSM_TRY(d2h_block(c, bands, sb, sB, eb, A->nb));
    if (U) SSM_TRY(d2h_block(c, U, su, sU, eu, A->nb));
    if (V) SSM_TRY(d2h_block(c, V, sv, sV, eu, A->nb));
    SSM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int ssm_ab_generate(ssm_ab *A, uint64_t seed, int dist) {
    SSM_CHECK_ARG(A, "null argument");
    SSM_CHECK_ARG(dist == SSM_DIST_DD || dist == SSM_DIST_BC, "unknown distribution");
    DeviceGuard g(A->ctx->device);
    SSM_TRY(ab_own_v(A, false));
    return launch_ab_generate(A->ctx, A, seed, dist);
}

static int abqr_alloc(const ssm_ab *A, bool with_q, ssm_abqr **out) {
    ssm_ctx *c = A->ctx;
    auto *F = new ssm_abqr();
    F->ctx = c; F->ref.bind(c); F->n = A->n; F->l = A->l; F->u = A->u; F->r = A->r; F->nb = A->nb; F->ntiles = A->ntiles; F->np = A->np;
    int rc = dev_alloc(c, (size_t)(F->ntiles * F->np * (F->l + F->u + 1 + F->r) * 32), true, F->RU);
    if (!rc && with_q) rc = dev_alloc(c, (size_t)(F->ntiles * F->np * (F->l + 1) * 32), true, F->QT);
    if (rc) { delete F; return rc; }
    F->V = A->V;
    *out = F;
    return 0;
}

int ssm_ab_qr_cols(const ssm_ab *A, int ncols, ssm_abqr **out) {
    SSM_CHECK_ARG(A && out, "null argument");
    if (ncols < 0 || ncols > A->n) { set_error("ssm_ab_qr_cols: ncols = %d outside 0..n = %d", ncols, A->n); return SSM_E_ARG; }
    ssm_ctx *c = A->ctx;
    DeviceGuard g(c->device);
    ssm_abqr *F = *out;
    CrossOrder co(c, {F ? F->ctx : nullptr});
    if (F) {   // refactor into an existing object of the same shape (steady-state benchmarking, no allocation)
        if (F->n != A->n || F->l != A->l || F->u != A->u || F->r != A->r || F->nb != A->nb) { set_error("ssm_ab_qr: factor object has a different shape"); return SSM_E_SHAPE; }
        F->V = A->V;
        uses(F->RU, c); uses(F->QT, c);          // a factor object of another context is rewritten on A's stream
    } else {
        SSM_TRY(abqr_alloc(A, true, &F));
    }
    AbArgs a{A->AU->p, A->V->p, F->RU->p, F->QT->p, nullptr, nullptr, A->n, A->np, ncols};
    int rc = launch_ab_qr(c, A->l, A->u, A->r, A->ntiles, a, false, true);
    if (rc) { if (!*out) delete F; return rc; }
    F->ncols = ncols;
    *out = F;
    return 0;
}
// _almostbanded_qr!(A, tau, ncols) continued: F holds the first F->ncols reflectors of A (ssm_ab_qr_cols); form reflectors F->ncols .. ncols-1 in
// place.  The result is what ssm_ab_qr_cols(A, ncols) returns.  What adaptive QR does when it needs more columns of an operand it has already
// begun to factor (the reference runs :139-172 on the trailing view; here the sweep re-enters at column F->ncols from the factor records).
int ssm_abqr_continue(ssm_abqr *F, const ssm_ab *A, int ncols) {
    SSM_CHECK_ARG(F && A, "null argument");
    if (F->n != A->n || F->l != A->l || F->u != A->u || F->r != A->r || F->nb != A->nb) { set_error("ssm_abqr_continue: DimensionMismatch: factor object and operand have different shapes"); return SSM_E_SHAPE; }
    if (ncols < F->ncols || ncols > A->n) { set_error("ssm_abqr_continue: ncols = %d outside %d..n = %d", ncols, F->ncols, A->n); return SSM_E_ARG; }
    if (!F->QT) { set_error("ssm_abqr_continue: factor object keeps no reflectors"); return SSM_E_STATE; }
    if (ncols == F->ncols) return 0;
    ssm_ctx *c = F->ctx;
    uses(A->AU, c); uses(A->V, c);
    DeviceGuard g(c->device);
    CrossOrder co(c, {A->ctx});
    AbArgs a{A->AU->p, A->V->p, F->RU->p, F->QT->p, nullptr, nullptr, A->n, A->np, ncols};
    SSM_TRY(F->ncols == 0 ? launch_ab_qr(c, A->l, A->u, A->r, A->ntiles, a, false, true) : launch_ab_qr_continue(c, A->l, A->u, A->r, A->ntiles, a, F->ncols));
    F->V = A->V;
    F->ncols = ncols;
    return 0;
}
// Grow a partial factorisation with its operand (adaptive QR on a cached operand, AlmostBandedMatrix.jl:357-394 + :139-172 on the trailing view):
// F = ssm_ab_qr_cols(A_old, k0), A = the operand after ssm_ab_resize + ssm_ab_set_rows (the old n0 x n0 part unchanged).  Afterwards F has A's
// size and equals ssm_ab_qr_cols(A, k0) to rounding, so ssm_abqr_continue(F, A, ncols) goes on from it.  The records of the old rows move into
// larger storage; the entries of rows >= n0 - 2l - u that reach into the new columns are re-derived by applying the stored reflectors to those
// columns (launch_ab_grow_fix); the rows from k0 on pass through the sweep once more without new reflectors.
int ssm_abqr_resize(ssm_abqr *F, const ssm_ab *A) {
    SSM_CHECK_ARG(F && A, "null argument");
    if (F->l != A->l || F->u != A->u || F->r != A->r || F->nb != A->nb) { set_error("ssm_abqr_resize: DimensionMismatch: factor object and operand have different bandwidths, rank or batch size"); return SSM_E_SHAPE; }
    if (A->n < F->n) { set_error("ssm_abqr_resize: DimensionMismatch: cannot shrink a %d x %d factorisation to %d x %d", F->n, F->n, A->n, A->n); return SSM_E_SHAPE; }
    if (!F->QT) { set_error("ssm_abqr_resize: factor object keeps no reflectors"); return SSM_E_STATE; }
    if (A->n > F->n && F->ncols + F->l > F->n) {
        set_error("ssm_abqr_resize: %d reflectors of a %d x %d matrix with lower bandwidth %d: the last of them saw columns cut off by the old size and do not "
                  "belong to the grown matrix (factor at most n - l columns before growing, as adaptive QR does)", F->ncols, F->n, F->n, F->l);
        return SSM_E_STATE;
    }
    ssm_ctx *c = F->ctx;
    uses(A->AU, c); uses(A->V, c);
    if (A->n == F->n) { F->V = A->V; return 0; }
    DeviceGuard g(c->device);
    CrossOrder co(c, {A->ctx});
    const int n0 = F->n, k0 = F->ncols;
    DevBufPtr RU, QT;
    SSM_TRY(dev_alloc(c, (size_t)(A->ntiles * A->np * (F->l + F->u + 1 + F->r) * 32), true, RU));
    SSM_TRY(dev_alloc(c, (size_t)(A->ntiles * A->np * (F->l + 1) * 32), true, QT));
    SSM_TRY(copy_tile_rows(c, RU->p, A->np, F->RU->p, F->np, n0, F->l + F->u + 1 + F->r, F->ntiles));
    SSM_TRY(copy_tile_rows(c, QT->p, A->np, F->QT->p, F->np, n0, F->l + 1, F->ntiles));
    SSM_CUDA(cudaStreamSynchronize(c->stream));       // the old blocks are released below: nothing may still be reading them
    F->RU = RU; F->QT = QT; F->V = A->V; F->n = A->n; F->np = A->np;
    AbArgs a{A->AU->p, A->V->p, F->RU->p, F->QT->p, nullptr, nullptr, A->n, A->np, k0};
    if (k0 == 0) return launch_ab_qr(c, A->l, A->u, A->r, A->ntiles, a, false, true);
    SSM_TRY(launch_ab_grow_fix(c, A->l, A->u, A->r, A->ntiles, a, n0, k0));
    return launch_ab_qr_continue(c, A->l, A->u, A->r, A->ntiles, a, k0);
}
int ssm_ab_qr(const ssm_ab *A, ssm_abqr **out) {
    SSM_CHECK_ARG(A && out, "null argument");
    return ssm_ab_qr_cols(A, A->n - 1, out);     // min(m - 1, n) for a real square matrix (:137)
}
int ssm_abqr_destroy(ssm_abqr *F) { if (F) { obj_sync(F->ctx); delete F; } return 0; }

int ssm_abqr_get(const ssm_abqr *F, double *factors, int64_t sf, double *U, int64_t su, double *tau, int64_t st) {
    SSM_CHECK_ARG(F, "null argument");
    ssm_ctx *c = F->ctx;
    DeviceGuard g(c->device);
    const size_t ef = (size_t)(2 * F->l + F->u + 1) * F->n, eu = (size_t)F->n * F->r, et = (size_t)F->n;
    if (sf == 0) sf = (int64_t)ef;
    if (su == 0) su = (int64_t)eu;
    if (st == 0) st = (int64_t)et;
    const bool anyhost = (factors && !is_device_ptr(factors)) || (U && !is_device_ptr(U)) || (tau && !is_device_ptr(tau));
    if (!anyhost) return launch_abqr_unpack(c, F, factors, sf, U, su, tau, st);
    double *stg;
    SSM_TRY(staging_dev(c, (size_t)F->nb * (ef + eu + et), &stg));
    double *sF = stg, *sU = sF + (size_t)F->nb * ef, *sT = sU + (size_t)F->nb * eu;
    SSM_TRY(launch_abqr_unpack(c, F, factors ? sF : nullptr, (int64_t)ef, U ? sU : nullptr, (int64_t)eu, tau ? sT : nullptr, (int64_t)et));
    if (factors) SSM_TRY(d2h_block(c, factors, sf, sF, ef, F->nb));
    if (U) SSM_TRY(d2h_block(c, U, su, sU, eu, F->nb));
    if (tau) SSM_TRY(d2h_block(c, tau, st, sT, et, F->nb));
    SSM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// A factor object from the reference's own fields (F.factors = widened bands + fill (U, V), F.tau; AlmostBandedMatrix.jl:174-185): the
// inverse of ssm_abqr_get.  What the glue calls for a QR it did not make itself (a copy, a deserialised factorisation) before ldiv!.
int ssm_abqr_set(ssm_ctx *c, int n, int l, int u, int r, int64_t nb, const double *factors, int64_t sf, const double *U, int64_t su,
                 const double *V, int64_t sv, const double *tau, int64_t st, int ncols, ssm_abqr **out) {
    SSM_CHECK_ARG(c && out && factors && tau && (r == 0 || (U && V)), "null argument");
    SSM_CHECK_ARG(n >= 1 && nb >= 1 && l >= 0 && u >= 0 && r >= 0, "bad sizes");
    if (l > SSM_MAX_BW || u > SSM_MAX_BW || r > SSM_MAX_RANK) { set_error("(l,u,r)=(%d,%d,%d) exceeds compiled limits (%d,%d,%d)", l, u, r, SSM_MAX_BW, SSM_MAX_BW, SSM_MAX_RANK); return SSM_E_UNSUPPORTED; }
    if (ncols < 0) ncols = n - 1;
    if (ncols > n) { set_error("ssm_abqr_set: ncols = %d outside 0..n = %d", ncols, n); return SSM_E_ARG; }
    const size_t ef = (size_t)(2 * l + u + 1) * n, eu = (size_t)n * r, et = (size_t)n;
    if (sf == 0) sf = (int64_t)ef;
    if (su == 0) su = (int64_t)eu;
    if (sv == 0) sv = (int64_t)eu;
    if (st == 0) st = (int64_t)et;
    if ((size_t)sf < ef || (size_t)su < eu || (size_t)sv < eu || (size_t)st < et) { set_error("ssm_abqr_set: stride smaller than one system"); return SSM_E_SHAPE; }
    DeviceGuard g(c->device);
    auto *F = new ssm_abqr();
    F->ctx = c; F->ref.bind(c); F->n = n; F->l = l; F->u = u; F->r = r; F->nb = nb; F->ntiles = ntiles_of(nb); F->np = (int64_t)n + PAD; F->ncols = ncols;
    int rc = dev_alloc(c, (size_t)(F->ntiles * F->np * (l + u + 1 + r) * 32), true, F->RU);
    if (!rc) rc = dev_alloc(c, (size_t)(F->ntiles * F->np * (l + 1) * 32), true, F->QT);
    if (!rc) rc = dev_alloc(c, (size_t)(F->ntiles * F->np * (r > 0 ? r : 1) * 32), true, F->V);
    if (rc) { delete F; return rc; }
    const bool dev = is_device_ptr(factors) && is_device_ptr(tau) && (r == 0 || (is_device_ptr(U) && is_device_ptr(V)));
    if (dev) rc = launch_abqr_pack(c, F, factors, sf, U, su, V, sv, tau, st);
    else {
        double *stg;
        rc = staging_dev(c, (size_t)nb * (ef + 2 * eu + et), &stg);
        double *sF = stg, *sU = sF + (size_t)nb * ef, *sV = sU + (size_t)nb * eu, *sT = sV + (size_t)nb * eu;
        if (!rc) rc = h2d_block(c, sF, factors, sf, ef, nb);
        if (!rc && r > 0) rc = h2d_block(c, sU, U, su, eu, nb);
        if (!rc && r > 0) rc = h2d_block(c, sV, V, sv, eu, nb);
        if (!rc) rc = h2d_block(c, sT, tau, st, et, nb);
        if (!rc) rc = launch_abqr_pack(c, F, sF, (int64_t)ef, sU, (int64_t)eu, sV, (int64_t)eu, sT, (int64_t)et);
        if (!rc) { cudaError_t e = cudaStreamSynchronize(c->stream); if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__); }
    }
    if (rc) { delete F; return rc; }
    *out = F;
    return 0;
}

static int check_vec(const char *fn, int n, int64_t nb, const ssm_ctx *c, const ssm_vec *b) {
    if (!b) { set_error("%s: null vector", fn); return SSM_E_ARG; }
    if (b->n != n || b->nb != nb) { set_error("%s: DimensionMismatch: vector batch is %lld x %d, matrix batch is %lld x %d", fn, (long long)b->nb, b->n, (long long)nb, n); return SSM_E_SHAPE; }
    if (b->ctx->device != c->device) { set_error("%s: operands live on different devices", fn); return SSM_E_ARG; }
    uses(b->d, c);              // a vector of another context is used on this one's stream: its block leaves the free-list protocol
    return 0;
}
// F's kernels run on F->ctx; its V (and, after a refactorisation from another context, its own records) may belong elsewhere
static void abqr_uses(const ssm_abqr *F) { uses(F->V, F->ctx); uses(F->RU, F->ctx); uses(F->QT, F->ctx); }
// The solve / apply entry points need the COMPLETE factorisation (n - 1 reflectors, AlmostBandedMatrix.jl:137): with fewer, "R" is not triangular.
static int abqr_complete(const char *fn, const ssm_abqr *F) {
    if (F->ncols < F->n - 1) { set_error("%s: the factor object holds %d of %d reflectors (ssm_ab_qr_cols / ssm_abqr_continue): finish it first", fn, F->ncols, F->n - 1); return SSM_E_STATE; }
    return 0;
}

int ssm_abqr_lmul_qt(const ssm_abqr *F, ssm_vec *b) {
    SSM_CHECK_ARG(F, "null argument");
    SSM_TRY(check_vec(__func__, F->n, F->nb, F->ctx, b));
    if (!F->QT) { set_error("factor object was created without reflectors"); return SSM_E_STATE; }
    abqr_uses(F);                  // a partial factorisation is fine here: the reflectors not yet formed have tau = 0
    DeviceGuard g(F->ctx->device);
    CrossOrder co(F->ctx, {b->ctx});
    AbSolveArgs a{F->RU->p, F->QT->p, F->V->p, b->d->p, b->d->p, F->n, F->np, 0};
    return launch_ab_qt(F->ctx, F->l, F->u, F->r, F->ntiles, a, true);
}
int ssm_abqr_lmul_q(const ssm_abqr *F, ssm_vec *b) {
    SSM_CHECK_ARG(F, "null argument");
    SSM_TRY(check_vec(__func__, F->n, F->nb, F->ctx, b));
    if (!F->QT) { set_error("factor object was created without reflectors"); return SSM_E_STATE; }
    abqr_uses(F);                  // a partial factorisation is fine here: the reflectors not yet formed have tau = 0
    DeviceGuard g(F->ctx->device);
    CrossOrder co(F->ctx, {b->ctx});
    AbSolveArgs a{F->RU->p, F->QT->p, F->V->p, b->d->p, b->d->p, F->n, F->np, 0};
    return launch_ab_qt(F->ctx, F->l, F->u, F->r, F->ntiles, a, false);
}
int ssm_abqr_upper_ldiv(const ssm_abqr *F, ssm_vec *b) {
    SSM_CHECK_ARG(F, "null argument");
    SSM_TRY(check_vec(__func__, F->n, F->nb, F->ctx, b));
    SSM_TRY(abqr_complete(__func__, F));
    abqr_uses(F);
    DeviceGuard g(F->ctx->device);
    CrossOrder co(F->ctx, {b->ctx});
    AbSolveArgs a{F->RU->p, F->QT ? F->QT->p : nullptr, F->V->p, b->d->p, b->d->p, F->n, F->np, 0};
    return launch_ab_backsolve(F->ctx, F->l, F->u, F->r, F->ntiles, a);
}
int ssm_abqr_solve(const ssm_abqr *F, const ssm_vec *b, ssm_vec *x) {
    SSM_CHECK_ARG(F, "null argument");
    SSM_TRY(check_vec(__func__, F->n, F->nb, F->ctx, b));
    SSM_TRY(check_vec(__func__, F->n, F->nb, F->ctx, x));
    if (!F->QT) { set_error("factor object was created without reflectors"); return SSM_E_STATE; }
    SSM_TRY(abqr_complete(__func__, F));
    abqr_uses(F);
    DeviceGuard g(F->ctx->device);
    CrossOrder co(F->ctx, {b->ctx, x->ctx});
    AbSolveArgs q{F->RU->p, F->QT->p, F->V->p, b->d->p, x->d->p, F->n, F->np, 0};        // x <- Q'b
    SSM_TRY(launch_ab_qt(F->ctx, F->l, F->u, F->r, F->ntiles, q, true));
    AbSolveArgs s{F->RU->p, F->QT->p, F->V->p, x->d->p, x->d->p, F->n, F->np, 0};        // x <- R \ x
    return launch_ab_backsolve(F->ctx, F->l, F->u, F->r, F->ntiles, s);
}
int ssm_abqr_ldiv(const ssm_abqr *F, ssm_vec *b) { return ssm_abqr_solve(F, b, b); }

int ssm_ab_solve(const ssm_ab *A, const ssm_vec *b, ssm_vec *x) {
    SSM_CHECK_ARG(A, "null argument");
    SSM_TRY(check_vec(__func__, A->n, A->nb, A->ctx, b));
    SSM_TRY(check_vec(__func__, A->n, A->nb, A->ctx, x));
    ssm_ctx *c = A->ctx;
    DeviceGuard g(c->device);
    CrossOrder co(c, {b->ctx, x->ctx});
    // R round-trips through HBM (the back-substitution runs against the sweep); reflectors are not kept.
    if (!A->scratch) SSM_TRY(dev_alloc(c, (size_t)(A->ntiles * A->np * (A->l + A->u + 1 + A->r) * 32), true, A->scratch));
    DevBufPtr RU = A->scratch;
    AbArgs a{A->AU->p, A->V->p, RU->p, nullptr, b->d->p, x->d->p, A->n, A->np, A->n - 1};
    SSM_TRY(launch_ab_qr(c, A->l, A->u, A->r, A->ntiles, a, true, false));
    AbSolveArgs s{RU->p, nullptr, A->V->p, x->d->p, x->d->p, A->n, A->np, 0};
    return launch_ab_backsolve(c, A->l, A->u, A->r, A->ntiles, s);
}

int ssm_ab_tri_ldiv(const ssm_ab *A, ssm_vec *b, int upper, int unit) {
    SSM_CHECK_ARG(A, "null argument");
    SSM_TRY(check_vec(__func__, A->n, A->nb, A->ctx, b));
    DeviceGuard g(A->ctx->device);
    CrossOrder co(A->ctx, {b->ctx});
    return launch_ab_tri(A->ctx, A, b->d->p, upper != 0, unit != 0);
}

int ssm_ab_residual(const ssm_ab *A, const ssm_vec *x, const ssm_vec *b, double *relres) {
    SSM_CHECK_ARG(A && relres, "null argument");
    SSM_TRY(check_vec(__func__, A->n, A->nb, A->ctx, x));
    SSM_TRY(check_vec(__func__, A->n, A->nb, A->ctx, b));
    ssm_ctx *c = A->ctx;
    DeviceGuard g(c->device);
    CrossOrder co(c, {x->ctx, b->ctx});
    if (is_device_ptr(relres)) return launch_ab_residual(c, A, x, b, relres);
    double *stg;
    SSM_TRY(staging_dev(c, (size_t)A->ntiles * 32, &stg));
    SSM_TRY(launch_ab_residual(c, A, x, b, stg));
    SSM_CUDA(cudaMemcpyAsync(relres, stg, (size_t)A->nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SSM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"
