// host_path.cu — ssm_ab_solve_host: the end-to-end `A \ b` for operands that live in HOST memory in the
// reference layout.  The batch is cut into chunks of systems; each chunk flows through
//     H2D (bands,U,V,b) -> pack (K4) -> qr with fused Q'b (K1) -> back-substitution (K3) -> unpack x -> D2H
// on one of NSLOT streams, so the PCIe copies of chunk c+1 overlap the kernels of chunk c and the D2H of
// chunk c-1.  Replaces, per system: `A \ b` = qr(A) then ldiv! (AlmostBandedMatrix.jl:125-133,187-192).
#include <algorithm>
#include "ssm_internal.cuh"

namespace ssm {

struct HostPipe {
    static constexpr int NSLOT = 3;
    int device = 0, n = 0, l = 0, u = 0, r = 0;
    int64_t ch = 0;
    ssm_ctx *cx[NSLOT] = {nullptr, nullptr, nullptr};
    ssm_ab *A[NSLOT] = {nullptr, nullptr, nullptr};
    ssm_vec *v[NSLOT] = {nullptr, nullptr, nullptr};
    DevBufPtr stg[NSLOT], ru[NSLOT];
    void release() {
        for (int s = 0; s < NSLOT; ++s) {
            if (A[s]) ssm_ab_destroy(A[s]);
            if (v[s]) ssm_vec_destroy(v[s]);
            stg[s].reset(); ru[s].reset();
            if (cx[s]) ssm_ctx_destroy(cx[s]);
            A[s] = nullptr; v[s] = nullptr; cx[s] = nullptr;
        }
    }
};

}  // namespace ssm

using namespace ssm;

// one pipeline per (ctx, shape); rebuilt when the shape changes
static HostPipe *get_pipe(ssm_ctx *c, int n, int l, int u, int r, int64_t nb, int *rc) {
    HostPipe *hp = reinterpret_cast<HostPipe *>(c->hostpipe);
    const int64_t want = std::min<int64_t>(8192, std::max<int64_t>(TILE, (((nb + HostPipe::NSLOT - 1) / HostPipe::NSLOT + TILE - 1) / TILE) * TILE));
    if (hp && hp->n == n && hp->l == l && hp->u == u && hp->r == r && hp->ch == want) { *rc = 0; return hp; }
    if (hp) { hp->release(); delete hp; c->hostpipe = nullptr; }
    hp = new HostPipe();
    hp->device = c->device; hp->n = n; hp->l = l; hp->u = u; hp->r = r; hp->ch = want;
    const size_t eb = (size_t)(l + u + 1) * n, eu = (size_t)n * r;
    for (int s = 0; s < HostPipe::NSLOT; ++s) {
        if ((*rc = ssm_ctx_create(c->device, &hp->cx[s]))) break;
        hp->cx[s]->ab_variant = c->ab_variant; hp->cx[s]->pipeline_stages = c->pipeline_stages;
        if ((*rc = ssm_ab_create(hp->cx[s], n, l, u, r, hp->ch, &hp->A[s]))) break;
        if ((*rc = ssm_vec_create(hp->cx[s], n, hp->ch, &hp->v[s]))) break;
        if ((*rc = dev_alloc(hp->cx[s], (size_t)hp->ch * (eb + 2 * eu + n), false, hp->stg[s]))) break;
        if ((*rc = dev_alloc(hp->cx[s], (size_t)(hp->A[s]->ntiles * hp->A[s]->np * (l + u + 1 + r) * 32), true, hp->ru[s]))) break;
    }
    if (*rc) { hp->release(); delete hp; return nullptr; }
    c->hostpipe = hp;
    return hp;
}

void ssm::hostpipe_free(ssm_ctx *c) {
    HostPipe *hp = reinterpret_cast<HostPipe *>(c->hostpipe);
    if (hp) { hp->release(); delete hp; c->hostpipe = nullptr; }
}

extern "C" int ssm_ab_solve_host(ssm_ctx *c, int n, int l, int u, int r, int64_t nb, const double *bands, const double *U,
                                 const double *V, const double *b, double *x) {
    if (!c || !bands || !b || !x || (r > 0 && (!U || !V))) { set_error("ssm_ab_solve_host: null argument"); return SSM_E_ARG; }
    if (n < 1 || nb < 1 || l < 0 || u < 0 || r < 0) { set_error("ssm_ab_solve_host: bad sizes"); return SSM_E_ARG; }
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(c->device);
    int rc = 0;
    HostPipe *hp = get_pipe(c, n, l, u, r, nb, &rc);
    if (!hp) { cudaSetDevice(prev); return rc; }
    const size_t eb = (size_t)(l + u + 1) * n, eu = (size_t)n * r;
    int64_t launches = 0;
    int slot = 0;
    for (int64_t s0 = 0; s0 < nb && !rc; s0 += hp->ch, slot = (slot + 1) % HostPipe::NSLOT) {
        const int64_t cnt = std::min(hp->ch, nb - s0);
        ssm_ctx *cx = hp->cx[slot];
        ssm_ab *A = hp->A[slot];
        ssm_vec *v = hp->v[slot];
        const int64_t l0 = cx->launches;
        double *sB = hp->stg[slot]->p, *sU = sB + (size_t)hp->ch * eb, *sV = sU + (size_t)hp->ch * eu, *sR = sV + (size_t)hp->ch * eu;
        cudaError_t e = cudaMemcpyAsync(sB, bands + (size_t)s0 * eb, (size_t)cnt * eb * 8, cudaMemcpyDefault, cx->stream);
        if (e == cudaSuccess && r > 0) e = cudaMemcpyAsync(sU, U + (size_t)s0 * eu, (size_t)cnt * eu * 8, cudaMemcpyDefault, cx->stream);
        if (e == cudaSuccess && r > 0) e = cudaMemcpyAsync(sV, V + (size_t)s0 * eu, (size_t)cnt * eu * 8, cudaMemcpyDefault, cx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(sR, b + (size_t)s0 * n, (size_t)cnt * n * 8, cudaMemcpyDefault, cx->stream);
        if (e != cudaSuccess) { rc = cuda_fail(e, "H2D chunk copy", __FILE__, __LINE__); break; }
        // the slot's batch objects describe `ch` systems; a short last chunk packs identity systems behind it
        A->nb = cnt; v->nb = cnt;
        const int64_t tiles = ntiles_of(cnt);
        if ((rc = launch_ab_pack(cx, A, sB, (int64_t)eb, sU, (int64_t)eu, sV, (int64_t)eu, 0, cnt))) break;
        if ((rc = launch_vec_pack(cx, v, sR, n, 0, cnt))) break;
        AbArgs a{A->AU->p, A->V->p, hp->ru[slot]->p, nullptr, v->d->p, v->d->p, n, A->np, n - 1};
        if ((rc = launch_ab_qr(cx, l, u, r, tiles, a, true, false))) break;
        AbSolveArgs s{hp->ru[slot]->p, nullptr, A->V->p, v->d->p, v->d->p, n, A->np, 0};
        if ((rc = launch_ab_backsolve(cx, l, u, r, tiles, s))) break;
        if ((rc = launch_vec_unpack(cx, v, sR, n, 0, cnt))) break;
        e = cudaMemcpyAsync(x + (size_t)s0 * n, sR, (size_t)cnt * n * 8, cudaMemcpyDefault, cx->stream);
        if (e != cudaSuccess) { rc = cuda_fail(e, "D2H chunk copy", __FILE__, __LINE__); break; }
        A->nb = hp->ch; v->nb = hp->ch;
        launches += cx->launches - l0;
    }
    for (int s = 0; s < HostPipe::NSLOT; ++s) {
        hp->A[s]->nb = hp->ch; hp->v[s]->nb = hp->ch;      // an error inside a short last chunk must not leave the cached pipeline truncated
        cudaError_t e = cudaStreamSynchronize(hp->cx[s]->stream);
        if (e != cudaSuccess && !rc) rc = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
    }
    c->launches += launches;
    cudaSetDevice(prev);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// Batch sharder: independent systems are split contiguously across the GPUs of one box, one host thread + one
// ssm_ctx (stream, chunk pipeline) per GPU, no collective — the host buffers ARE the gather (north_star).
// ---------------------------------------------------------------------------------------------------------
#include <mutex>
#include <thread>
#include <map>

namespace {
std::mutex g_shard_mu;
std::map<std::pair<int, int>, ssm_ctx *> g_shard_ctx;   // one lazily created context per (shard, device), kept for the process lifetime

ssm_ctx *shard_ctx(int shard, int device, int *rc) {
    std::lock_guard<std::mutex> lk(g_shard_mu);
    auto it = g_shard_ctx.find({shard, device});
    if (it != g_shard_ctx.end()) { *rc = 0; return it->second; }
    ssm_ctx *c = nullptr;
    *rc = ssm_ctx_create(device, &c);
    if (*rc) return nullptr;
    g_shard_ctx[{shard, device}] = c;
    return c;
}
}  // namespace

// systems [s0, s1) of a batch of nb go to shard g of G:  s0 = g*nb/G rounded down to a multiple of 32 (tile size)
extern "C" int ssm_shard_range(int64_t nb, int g, int G, int64_t *s0, int64_t *s1) {
    if (!s0 || !s1 || G < 1 || g < 0 || g >= G || nb < 0) { set_error("ssm_shard_range: bad arguments"); return SSM_E_ARG; }
    auto cut = [&](int k) -> int64_t { if (k >= G) return nb; int64_t c = (nb * k) / G; c -= c % TILE; return c; };
    *s0 = cut(g); *s1 = cut(g + 1);
    return 0;
}

extern "C" int ssm_ab_solve_host_sharded(int ndev, const int *devices, int n, int l, int u, int r, int64_t nb, const double *bands,
                                         const double *U, const double *V, const double *b, double *x) {
    if (ndev < 1 || !devices) { set_error("ssm_ab_solve_host_sharded: need at least one device"); return SSM_E_ARG; }
    if (!bands || !b || !x || (r > 0 && (!U || !V)) || n < 1 || nb < 1) { set_error("ssm_ab_solve_host_sharded: bad arguments"); return SSM_E_ARG; }
    const size_t eb = (size_t)(l + u + 1) * n, eu = (size_t)n * r;
    std::vector<int> rcs(ndev, 0);
    std::vector<std::string> errs(ndev);
    std::vector<std::thread> th;
    for (int g = 0; g < ndev; ++g) {
        th.emplace_back([&, g]() {
            int64_t s0 = 0, s1 = 0;
            ssm_shard_range(nb, g, ndev, &s0, &s1);
            if (s1 <= s0) return;
            int rc = 0;
            ssm_ctx *c = shard_ctx(g, devices[g], &rc);
            if (!rc) rc = ssm_ab_solve_host(c, n, l, u, r, s1 - s0, bands + (size_t)s0 * eb, r ? U + (size_t)s0 * eu : nullptr,
                                            r ? V + (size_t)s0 * eu : nullptr, b + (size_t)s0 * n, x + (size_t)s0 * n);
            rcs[g] = rc;
            if (rc) errs[g] = ssm_last_error();
        });
    }
    for (auto &t : th) t.join();
    for (int g = 0; g < ndev; ++g)
        if (rcs[g]) { set_error("shard %d (device %d): %s", g, devices[g], errs[g].c_str()); return rcs[g]; }
    return 0;
}
