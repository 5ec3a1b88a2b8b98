// stream_loader.cuh — how a warp pulls its tile's record streams out of HBM.
//
// A sweep consumes, per step g, one record from each of up to three streams (e.g. the entering row record,
// one V column, one right-hand-side value).  Each stream s is described by (base pointer of the tile for
// this lane, fields per record NFs, record offset OFFs): step g reads record  OFFs + DIR*g.
//
//   DirectLoader : ld.global into registers one step ahead (no shared memory).
//   RingLoader   : Blackwell/Hopper bulk-async copies (cp.async.bulk = TMA 1-D, SASS UBLKCP) issued by lane 0
//                  into a per-warp shared-memory ring of NSTAGE steps, completion tracked by one mbarrier per
//                  stage (complete_tx::bytes).  The warp reads its stage with conflict-free LDS.64 (lane-fastest
//                  records), then lane 0 re-arms the stage NSTAGE steps ahead.  Loads therefore run NSTAGE-1
//                  steps ahead of the arithmetic without costing registers.
#pragma once
#include <cstdint>

namespace ssm {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// Bounded wait: a lost completion must never hang the GPU box — trap after ~2^31 cycles instead.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > (1ll << 31)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}
// Orders this thread's (and, through a preceding barrier, its warp's / CTA's) generic-proxy accesses to shared memory before
// subsequent async-proxy operations.  REQUIRED between the last LDS of a stage and the bulk copy that re-arms it: without it the
// copy's data may land before the reads are performed.  (Found the hard way: a ring that passed every sequential test returned
// wrong rows as soon as another stream kept the GPU busy; see profiles/r1_measurements.md.)
__device__ __forceinline__ void fence_proxy_async() {
#ifndef SSM_TEST_WITHOUT_PROXY_FENCE     // only for checking that tests/test_gpu_concurrent.py catches the hazard
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- DirectLoader -------------------------------------------------------------------------------------
// STRs = distance between consecutive records of stream s, in fields (default: dense records, STRs = NFs)
template <int NF0, int NF1, int NF2, int DIR, int STR0 = NF0, int STR1 = NF1, int STR2 = NF2>
struct DirectLoader {
    const double *p0, *p1, *p2;   // lane-adjusted pointers to record (OFFs) of each stream
    int64_t nsteps;
    double n0[NF0 > 0 ? NF0 : 1], n1[NF1 > 0 ? NF1 : 1], n2[NF2 > 0 ? NF2 : 1];

    __device__ __forceinline__ void fetch() {
#pragma unroll
        for (int f = 0; f < NF0; ++f) n0[f] = p0[f * 32];
#pragma unroll
        for (int f = 0; f < NF1; ++f) n1[f] = p1[f * 32];
#pragma unroll
        for (int f = 0; f < NF2; ++f) n2[f] = p2[f * 32];
        p0 += DIR * STR0 * 32; p1 += DIR * STR1 * 32; p2 += DIR * STR2 * 32;
    }
    __device__ __forceinline__ void init(const double *b0, const double *b1, const double *b2, void *, int, int64_t total_steps) {
        p0 = b0; p1 = b1; p2 = b2; nsteps = total_steps;
        if (total_steps > 0) fetch();
    }
    // hands out step g's records and starts loading step g+1 (never touches a record past the last step)
    __device__ __forceinline__ void acquire(int64_t g, double *r0, double *r1, double *r2) {
#pragma unroll
        for (int f = 0; f < NF0; ++f) r0[f] = n0[f];
#pragma unroll
        for (int f = 0; f < NF1; ++f) r1[f] = n1[f];
#pragma unroll
        for (int f = 0; f < NF2; ++f) r2[f] = n2[f];
        if (g + 1 < nsteps) fetch();
    }
};

// ---- RingLoader ---------------------------------------------------------------------------------------
template <int NF0, int NF1, int NF2, int DIR, int STR0 = NF0, int STR1 = NF1, int STR2 = NF2>
struct RingLoader {
    static constexpr int STAGE_DOUBLES = (NF0 + NF1 + NF2) * 32;
    static constexpr uint32_t STAGE_BYTES = STAGE_DOUBLES * 8;
    const double *g0, *g1, *g2;   // tile base (NOT lane adjusted) of record OFFs of each stream
    double *ring;                 // this warp's ring
    uint32_t bar0;                // smem address of mbarrier[0]
    int nstage;
    int64_t nsteps;               // total steps that will be consumed
    int lane;

    static __host__ __device__ size_t smem_bytes(int nstage) { return (size_t)nstage * STAGE_BYTES + (size_t)nstage * 8; }

    int cur = 0;                  // stage holding the next step to be consumed
    uint32_t parity = 0;          // its mbarrier phase parity

    // arm stage s with step g (s == g mod nstage by construction; no division on the hot path)
    __device__ __forceinline__ void issue(int64_t g, int s) {
        const uint32_t bar = bar0 + 8u * s;
        double *dst = ring + (size_t)s * STAGE_DOUBLES;
        mbar_expect_tx(bar, STAGE_BYTES);
        if (NF0 > 0) bulk_g2s(smem_u32(dst), g0 + (int64_t)DIR * g * STR0 * 32, NF0 * 256, bar);
        if (NF1 > 0) bulk_g2s(smem_u32(dst + NF0 * 32), g1 + (int64_t)DIR * g * STR1 * 32, NF1 * 256, bar);
        if (NF2 > 0) bulk_g2s(smem_u32(dst + (NF0 + NF1) * 32), g2 + (int64_t)DIR * g * STR2 * 32, NF2 * 256, bar);
    }
    // b0..b2: lane-adjusted pointers as for DirectLoader (lane offset removed here)
    __device__ __forceinline__ void init(const double *b0, const double *b1, const double *b2, void *smem, int nst,
                                         int64_t total_steps) {
        lane = threadIdx.x & 31;
        g0 = b0 - lane; g1 = b1 - lane; g2 = b2 - lane;
        nstage = nst; nsteps = total_steps;
        ring = reinterpret_cast<double *>(smem);
        bar0 = smem_u32(reinterpret_cast<unsigned char *>(smem) + (size_t)nst * STAGE_BYTES);
        if (lane == 0) {
            for (int s = 0; s < nst; ++s) mbar_init(bar0 + 8u * s, 1);
            fence_barrier_init();
        }
        __syncwarp();
        if (lane == 0) {
            for (int g = 0; g < nst && g < total_steps; ++g) issue(g, g);
        }
    }
    // steps must be acquired in order g = 0, 1, 2, ...
    __device__ __forceinline__ void acquire(int64_t g, double *r0, double *r1, double *r2) {
        const int s = cur;
        mbar_wait(bar0 + 8u * s, parity);
        const double *src = ring + (size_t)s * STAGE_DOUBLES + lane;
#pragma unroll
        for (int f = 0; f < NF0; ++f) r0[f] = src[f * 32];
#pragma unroll
        for (int f = 0; f < NF1; ++f) r1[f] = src[(NF0 + f) * 32];
#pragma unroll
        for (int f = 0; f < NF2; ++f) r2[f] = src[(NF0 + NF1 + f) * 32];
        __syncwarp();   // every lane has its copy in registers before the stage is re-armed
        fence_proxy_async();   // all lanes: their reads of this stage are performed before the re-arming bulk copy may write it
        if (lane == 0 && g + nstage < nsteps) issue(g + nstage, s);
        if (++cur == nstage) { cur = 0; parity ^= 1u; }
    }
};

// dense-record aliases (usable as 4-parameter template template arguments)
template <int NF0, int NF1, int NF2, int DIR> using DirectLoaderD = DirectLoader<NF0, NF1, NF2, DIR>;
template <int NF0, int NF1, int NF2, int DIR> using RingLoaderD = RingLoader<NF0, NF1, NF2, DIR>;

// ---- GroupRing ----------------------------------------------------------------------------------------
// The loader for SMALL batches, where one warp per scheduler (or fewer) runs alone and every cycle spent on the loader sits
// on the sweep's serial chain.  Differences from RingLoader:
//   * a stage holds a GROUP of gm blocks of P consecutive steps (P = the sweep's unroll period); the rows of a group are
//     contiguous in the tile's record stream, so a group is ONE bulk copy per stream and ONE mbarrier wait per gm*P steps;
//   * the copies are issued by a dedicated producer warp (one thread), which waits on per-stage "empty" mbarriers — the
//     consumers never execute a copy instruction;
// Inside a block the reads are plain LDS at compile-time offsets from three pointers, so the compiler hoists them freely.
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// STRs > NFs: stream s is a field range of wider records (STRs fields apart); its rows are then copied one by one (still by the
// producer, still one mbarrier wait per group for the consumer).
template <int NF0, int NF1, int NF2, int DIR, int P, int STR0 = NF0, int STR1 = NF1, int STR2 = NF2>
struct GroupRing {
    static constexpr int RECD = (NF0 + NF1 + NF2) * 32;                 // doubles per step
    static constexpr size_t BLOCK_BYTES = (size_t)P * RECD * 8;
    static __host__ __device__ size_t smem_bytes(int nstage, int gm) { return (size_t)nstage * gm * BLOCK_BYTES + (size_t)nstage * 16; }

    // every thread of the CTA calls this once (contains a CTA barrier)
    static __device__ __forceinline__ void init_barriers(void *smem, int nstage, int gm, int nconsumer_warps) {
        const uint32_t bars = smem_u32(reinterpret_cast<unsigned char *>(smem) + (size_t)nstage * gm * BLOCK_BYTES);
        if (threadIdx.x == 0) {
            for (int s = 0; s < nstage; ++s) { mbar_init(bars + 16u * s, 1); mbar_init(bars + 16u * s + 8u, nconsumer_warps); }
            fence_barrier_init();
        }
        __syncthreads();
    }
    template <int NF, int STR>
    static __device__ __forceinline__ void copy_stream(double *dst, const double *g, int64_t lo, int rows, uint32_t bar) {
        if constexpr (NF > 0 && STR == NF) {
            bulk_g2s(smem_u32(dst), g + lo * NF * 32, (uint32_t)rows * NF * 256u, bar);
        } else if constexpr (NF > 0) {
            for (int r = 0; r < rows; ++r) bulk_g2s(smem_u32(dst + (size_t)r * NF * 32), g + (lo + r) * STR * 32, NF * 256u, bar);
        }
    }
    // producer thread: g0..g2 = tile base (not lane adjusted) of the record step 0 reads from each stream
    static __device__ __forceinline__ void produce(const double *g0, const double *g1, const double *g2, void *smem, int nstage,
                                                   int gm, int64_t nblocks) {
        double *ring = reinterpret_cast<double *>(smem);
        const int G = gm * P;
        const uint32_t bars = smem_u32(reinterpret_cast<unsigned char *>(smem) + (size_t)nstage * gm * BLOCK_BYTES);
        int s = 0; uint32_t use = 0;
        for (int64_t b0 = 0; b0 < nblocks; b0 += gm) {
#ifdef SSM_GROUP_NO_FENCE_P
            if (use > 0) { mbar_wait(bars + 16u * s + 8u, (use - 1) & 1u); }
#else
            if (use > 0) { mbar_wait(bars + 16u * s + 8u, (use - 1) & 1u); fence_proxy_async(); }
#endif
            const int rows = (int)((nblocks - b0 < gm ? nblocks - b0 : gm)) * P;
            const int64_t r0 = b0 * P;                                   // first step of the group
            const int64_t lo = DIR > 0 ? r0 : -(r0 + rows - 1);          // lowest record of the group, relative to step 0's
            const uint32_t bar = bars + 16u * s;
            double *dst = ring + (size_t)s * G * RECD;
            mbar_expect_tx(bar, (uint32_t)rows * RECD * 8u);
            copy_stream<NF0, STR0>(dst, g0, lo, rows, bar);
            copy_stream<NF1, STR1>(dst + (size_t)G * NF0 * 32, g1, lo, rows, bar);
            copy_stream<NF2, STR2>(dst + (size_t)G * (NF0 + NF1) * 32, g2, lo, rows, bar);
            if (++s == nstage) { s = 0; ++use; }
        }
    }

    // ---- consumer side ----
    double *ring; uint32_t bars; int nstage, gm, lane, col;
    int64_t blocks_left;
    int cur = 0, blk = 0, cb = 0; uint32_t parity = 0;
    const double *s0, *s1, *s2;
    __device__ __forceinline__ void attach(void *smem, int nst, int gm_, int64_t nblocks, int col_) {
        ring = reinterpret_cast<double *>(smem); nstage = nst; gm = gm_; blocks_left = nblocks; col = col_;
        lane = threadIdx.x & 31;
        bars = smem_u32(reinterpret_cast<unsigned char *>(smem) + (size_t)nst * gm_ * BLOCK_BYTES);
    }
    __device__ __forceinline__ void block_begin() {
        if (blk == 0) {
            mbar_wait(bars + 16u * cur, parity);
            cb = (int)(blocks_left < gm ? blocks_left : gm);
        }
        const int G = gm * P;
        const int row = DIR > 0 ? blk * P : cb * P - 1 - blk * P;
        const double *base = ring + (size_t)cur * G * RECD + col;
        s0 = base + row * NF0 * 32;
        s1 = base + G * NF0 * 32 + row * NF1 * 32;
        s2 = base + G * (NF0 + NF1) * 32 + row * NF2 * 32;
    }
    // t = step within the block, in consumption order (compile-time at the call site)
    __device__ __forceinline__ void read(const int t, double *r0, double *r1, double *r2) const {
#pragma unroll
        for (int f = 0; f < NF0; ++f) r0[f] = s0[(DIR * t * NF0 + f) * 32];
#pragma unroll
        for (int f = 0; f < NF1; ++f) r1[f] = s1[(DIR * t * NF1 + f) * 32];
#pragma unroll
        for (int f = 0; f < NF2; ++f) r2[f] = s2[(DIR * t * NF2 + f) * 32];
    }
    __device__ __forceinline__ void block_end() {
        if (++blk == cb) {
#ifndef SSM_GROUP_NO_FENCE_C
            fence_proxy_async();          // this lane's reads of the stage precede the producer's next bulk copy into it
#endif
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 16u * cur + 8u);
            blk = 0; blocks_left -= cb;
            if (++cur == nstage) { cur = 0; parity ^= 1u; }
        }
    }
};

// ---- WindowRing ---------------------------------------------------------------------------------------
// A sliding window of whole records that stay in shared memory while the sweep needs them (the BPS sweep looks
// M rows back and L rows ahead).  Record k of the sequence lives in slot k mod NST (NST a power of two) from
// the moment its bulk copy lands until release(k) re-arms the slot with record k+NST.  Lanes read fields in
// place: rec(k)[f*32] is conflict-free (lane-fastest records).
// NW = 1: the ring belongs to one warp.  NW = 2: two warps of one CTA share it — each computes 16 of the tile's 32 systems (set
// `col`), thread 0 issues the copies, and the slot is re-armed after a CTA barrier (kernels that are latency bound at one warp per
// tile double their warps in flight this way without moving a byte more).  NST need not be a power of two.
// STRIDE: fields between consecutive records in global memory (> NF when only a field range of each record is wanted)
template <int NF, int NST, int NW = 1, int STRIDE = NF>
struct WindowRing {
    static constexpr uint32_t REC_BYTES = NF * 256;
    const double *g;      // tile base of record 0 of the sequence (not lane adjusted)
    double *ring;
    uint32_t bar0;
    int col;              // lane column of the tile this thread reads
    bool issuer;
    int64_t nseq;
    static constexpr __host__ __device__ size_t smem_bytes() { return (size_t)NST * REC_BYTES + (size_t)NST * 8; }
    __device__ __forceinline__ void issue(int64_t k) {
        const int s = (int)(k % NST);
        const uint32_t bar = bar0 + 8u * s;
        mbar_expect_tx(bar, REC_BYTES);
        bulk_g2s(smem_u32(ring + (size_t)s * NF * 32), g + k * STRIDE * 32, REC_BYTES, bar);
    }
    __device__ __forceinline__ void sync() { if (NW == 1) __syncwarp(); else __syncthreads(); }
    __device__ __forceinline__ void init(const double *tile_first_record, void *smem, int64_t total) {
        col = threadIdx.x & 31;
        issuer = threadIdx.x == 0;
        g = tile_first_record; nseq = total;
        ring = reinterpret_cast<double *>(smem);
        bar0 = smem_u32(reinterpret_cast<unsigned char *>(smem) + (size_t)NST * REC_BYTES);
        if (issuer) {
            for (int s = 0; s < NST; ++s) mbar_init(bar0 + 8u * s, 1);
            fence_barrier_init();
        }
        sync();
        if (issuer) for (int64_t k = 0; k < NST && k < total; ++k) issue(k);
    }
    __device__ __forceinline__ void wait(int64_t k) { mbar_wait(bar0 + 8u * (uint32_t)(k % NST), (uint32_t)((k / NST) & 1)); }
    // the same three operations for callers that track (slot, parity) of a record themselves — no division on the hot path
    __device__ __forceinline__ void wait_slot(int slot, uint32_t parity) { mbar_wait(bar0 + 8u * (uint32_t)slot, parity); }
    __device__ __forceinline__ const double *slot_ptr(int slot) const { return ring + slot * (NF * 32) + col; }
    __device__ __forceinline__ void release_slot(int64_t k, int slot) {      // record k lives in `slot`
        sync();
        fence_proxy_async();
        if (issuer && k + NST < nseq) {
            const uint32_t bar = bar0 + 8u * slot;
            mbar_expect_tx(bar, REC_BYTES);
            bulk_g2s(smem_u32(ring + (size_t)slot * NF * 32), g + (k + NST) * STRIDE * 32, REC_BYTES, bar);
        }
    }
    __device__ __forceinline__ const double *rec(int64_t k) const { return ring + (size_t)(k % NST) * NF * 32 + col; }
    __device__ __forceinline__ void release(int64_t k) {
        sync();
        fence_proxy_async();
        if (issuer && k + NST < nseq) issue(k + NST);
    }
};

}  // namespace ssm
