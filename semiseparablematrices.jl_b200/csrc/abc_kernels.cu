// abc_kernels.cu — K10: ONE large AlmostBandedMatrix system, chunked along the n axis (BASELINE config 4:
// n = 2^24, l = u = 4, r = 2).  Replaces, for a single system whose sequential sweep (AlmostBandedMatrix.jl:144-170,
// :224-238) cannot fill a GPU, the whole `A \ b` (AlmostBandedMatrix.jl:125-133, 187-199) by a stable partitioned
// orthogonal factorisation that returns the same SOLUTION (the factors are not the reference's):
//
//   stage 1  abc_qr      Householder elimination of each row chunk's interior columns, 8 lanes per chunk and 4 chunks per warp.
//                        Lanes hold the COLUMNS of the active block (window | generator | separator coefficients | tail
//                        coefficients | rhs), reflectors are broadcast with shuffles; rows and window columns rotate through fixed
//                        slots/lanes so nothing moves when the window slides.  The entering operand rows and V rows arrive through
//                        a per-warp bulk-async ring (TMA 1-D copies), so the sweep's only global load is the rhs value.
//                        Large systems of the shapes (l+u, r) = (8 | 6 | 4, 2): abc_qr_lpc, ONE LANE per chunk — four warps per 32 chunks
//                        share a chunk's 22 columns (reflector | rest of the panel | trailing columns x 2), coupled by shared-memory
//                        rings, no shuffles, a rolled step loop; entering rows from a chunk-interleaved copy of the operands
//                        (abc_interleave), records interleaved over the tile of 32 chunks.  Wide shapes otherwise: 4 lanes per chunk.
//   stage 2  abc_merge   cyclic reduction of the block-bidiagonal separator system F_c z_c + G_c z_{c+1} = h_c by QR of
//                        [G_a; F_b], one warp per merge, log2(P) launches; abc_root solves the last block; abc_expand
//                        recovers the eliminated separators top-down.
//   stage 3  abc_bs      back-substitution of the interior unknowns.  Many chunks (>= ssm_ctx option "abc_tiled_min_chunks"):
//                        abc_bs_tiled, ONE LANE PER CHUNK — stage 1 writes its R-row records interleaved over mini-tiles of 4
//                        chunks (one stage-1 warp), a stage-3 warp streams eight mini-tiles through a bulk-async ring and every
//                        lane runs the plain scalar recurrence of its own chunk.  Few chunks: abc_bs, one warp
//                        per chunk with lanes = rows of a 32-row block (row-major records).
//
// The numpy restatement tests/emul/sof_reference.py documents the algebra and is the CPU check of this file.
// Layout (row-major, one system): REC[n+pad][lp+1+r] = A[i, i-l .. i+u] | U[i, :];  V[n+pad][r];  lp = l + u.
#include <algorithm>
#include "../../include/ssm_rng.h"
#include "fastmath.cuh"
#include "ssm_internal.cuh"
#include "stream_loader.cuh"

namespace ssm {

constexpr unsigned FULL = 0xffffffffu;
// predicated 8-byte store (no branch: a conditional block is a scheduling fence in the unrolled sweeps)
__device__ __forceinline__ void st_global_if(double *p, const double v, const bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.global.f64 [%0], %1;\n\t}" ::"l"(p), "d"(v), "r"((int)on) : "memory");
}
constexpr int ABC_PAD = 512;    // identity rows behind the system: room for the regular chunk layout (abc_layout) and read-ahead
constexpr int ABC_CPW = 4;      // chunks per warp in stage 1: chunk lengths change only at multiples of this

struct AbcArgs {
    const double *REC, *V, *b;   // operands
    double *FR;                  // R-row records, canonical column order: [n][NC] (row-major) or tiled (see npf)
    double *E0;                  // [P][W][2W+1] level-0 reduced blocks [F | G | h]
    const double *Z;             // [(P+1)][W] separator solution (stage 3)
    double *x;
    int64_t n; int l, u; int P;
    int m0, jx, pw;              // chunk c has m0 + pw*(c < jx) rows: every chunk length is == lp (mod lp+1), see abc_layout()
    int64_t npf;                 // tiled FR: records per mini-tile of 4 chunks, FR[c / 4][npf][NC][c % 4]
    int *smc;                    // lane-per-chunk stage 1: one arrival counter per SM (never reset: only its low bits are used)
    const double *RI, *VI;       // lane-per-chunk stage 1: chunk-interleaved entering rows / V rows (abc_interleave_kernel), nrow_i steps per tile
    int nrow_i;
};

// Regular chunks: the system is padded with < lp+1 identity rows to n' = P*m0 + jx*pw so that every chunk eliminates a
// multiple of pw = lp+1 columns (the sweep is unrolled by pw; no remainder code, no divergence around the shuffles).
__host__ __device__ __forceinline__ int64_t chunk_start(const AbcArgs &a, int c) { return (int64_t)c * a.m0 + (int64_t)a.pw * (c < a.jx ? c : a.jx); }


// ---------------------------------------------------------------------------------------------------------------
// stage 1
// ---------------------------------------------------------------------------------------------------------------
template <int LP, int R>
struct AbcDims {
    static constexpr int PW = LP + 1;                 // window columns = band-row slots
    static constexpr int NR = LP + 1 + R;             // active rows (band slots + tail rows)
    static constexpr int NC = 2 * LP + 2 * R + 2;     // active columns = lanes in use
    static constexpr int G0 = PW, C0 = PW + R, T0 = PW + R + LP, B0 = NC - 1;
    static constexpr int NF = PW + R;                 // fields of an operand row record
    static constexpr int W = LP + R;                  // separator block size
    static_assert(NC <= 32, "active block must fit one warp");
    static_assert(3 * W + 1 <= 32, "merge block must fit one warp");
};

// GS lanes cooperate on one chunk, so a warp sweeps 32/GS chunks at once: lane gl of a group holds the columns
// col = slot*GS + gl, slot = 0..CS-1, of the chunk's active block.  (GS = 32 is one chunk per warp; GS = 8 cuts the shuffle and
// redundant-reflector cost per chunk by 4 — every shuffle instruction serves four chunks.)  All chunks of one warp have the
// same length (abc_layout makes jx a multiple of 32/GS), so the groups stay in lock step.
// TILED: the R-row records are interleaved over the FOUR chunks this warp sweeps, FR[c / 4][k][field][c % 4]: one store
// instruction (8 lanes = 8 fields, 4 chunks each) then covers 256 contiguous bytes — 2-3 cache lines against ~5 for row-major
// records and 8 for 32-wide tiles — which matters because the LSU pipe (shuffles + global accesses) is this kernel's busiest unit.
// Shared memory of stage 1 (dynamic).  Each warp owns one region: a ring of NSTG stages, a stage = the PW operand rows and the PW
// V rows that enter during one block of PW steps, for each of the warp's chunks (one bulk-async copy per chunk and stream; the
// copies start at the 16-byte boundary at or below the first row, hence the 16 spare bytes and a per-block shift of 0 or 1
// doubles).  The rhs values of a block (caller's array: no padding to over-read) come by 8-byte cp.async with zero fill past the end.
// The ring part of the region is the transposition buffer of the epilogue once the ring has drained.
template <int LP, int R, int GS>
struct AbcQrSmem {
    using D = AbcDims<LP, R>;
    static constexpr int CPW = 32 / GS, CS = (D::NC + GS - 1) / GS, NSTG = GS <= 4 ? 2 : 3;   // (a stage lasts PW steps; at GS = 4 two CTAs per SM need the room)
    static constexpr int CBR = ((D::PW * D::NF * 8 + 15) / 16) * 16 + 16;                     // bytes per chunk and stage: operand rows
    static constexpr int CBV = R > 0 ? ((D::PW * R * 8 + 15) / 16) * 16 + 16 : 0;             // ... V rows
    static constexpr int SGB = CPW * (CBR + CBV);                                            // bytes per stage
    static constexpr int EPI = CS * D::NR * 32 * 8;                                          // epilogue buffer of one warp
    static constexpr int PWE = D::PW + (D::PW & 1);                                          // rhs values per chunk and stage (even count)
    static constexpr int BOFF = (NSTG * SGB > EPI ? NSTG * SGB : EPI);                       // rhs buffer [NSTG][CPW][PWE] behind the ring
    static constexpr int WREG = ((BOFF + NSTG * CPW * PWE * 8 + 127) / 128) * 128;
    static constexpr int BARS = 4 * WREG, ZERO = BARS + 4 * NSTG * 8, TOTAL = ZERO + ((D::PW * D::NF * 8 + 15) / 16) * 16;   // ZERO: one block of zero rows
};

template <int LP, int R, int GS, bool TILED>
__global__ void __launch_bounds__(128) abc_qr_kernel(const AbcArgs a) {
    // BF: branch-free step (reflector scalars masked instead of selected under a branch, predicated record stores): with the six
    // column slots of GS = 4 the scheduler then has independent arithmetic to put beside the scalar chain; at GS = 8 the extra
    // registers cost the third CTA per SM and the branches stay (measured: profiles/r2_measurements.md)
    constexpr bool BF = GS <= 4;
    using D = AbcDims<LP, R>;
    using SM = AbcQrSmem<LP, R, GS>;
    constexpr int PW = D::PW, NR = D::NR, NC = D::NC, G0 = D::G0, C0 = D::C0, T0 = D::T0, B0 = D::B0, NF = D::NF, W = D::W;
    constexpr int CS = (NC + GS - 1) / GS, CPW = 32 / GS, RQ = R > 0 ? R : 1;
    constexpr int FSTR = TILED ? 4 : 1, RSTR = TILED ? NC * 4 : NC;     // field / record strides of FR in doubles
    constexpr int NSTG = SM::NSTG;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, gl = lane % GS, grp = lane / GS;
    unsigned char *wreg = smem + wib * SM::WREG;                       // this warp's ring / epilogue region
    double (*stage)[NR][32] = reinterpret_cast<double (*)[NR][32]>(wreg);
    int c = (blockIdx.x * 4 + wib) * CPW + grp;
    if ((blockIdx.x * 4 + wib) * CPW >= a.P) return;
    const bool live = c < a.P;                                // a partial last warp repeats chunk P-1 in its spare groups (no stores)
    if (!live) c = a.P - 1;
    const int64_t r0 = chunk_start(a, c);
    const int m = a.m0 + (c < a.jx ? PW : 0), nsteps = m - LP;      // a multiple of PW, the same for every group of the warp
    const int u = a.u;
    const double *rec0 = a.REC + r0 * NF;
    const double *v0 = a.V + (r0 + u) * RQ;                   // V row of local column 0
    const double *b0 = a.b + r0;
    const int64_t brows = a.n - r0;                           // rows of b that exist from r0 on (pad rows have b = 0)
    // ---- ring of entering rows: lanes 0..CPW-1 copy the operand rows of the warp's chunks, lanes CPW..2CPW-1 their V rows ----
    const uint32_t bar0 = smem_u32(smem + SM::BARS + wib * NSTG * 8);
    double *zero_s = reinterpret_cast<double *>(smem + SM::ZERO);
    const int nblk_ring = nsteps / PW + 1;                    // the row fetched ahead at the last step belongs to one more block
    const double *csrc = nullptr;                             // issuing lanes: first row / V row of block 0 of "their" chunk
    {
        int cc = (blockIdx.x * 4 + wib) * CPW + (lane % CPW);
        if (cc >= a.P) cc = a.P - 1;
        const int64_t rc = chunk_start(a, cc);
        csrc = lane < CPW ? a.REC + (rc + LP) * NF : a.V + (rc + u + LP + 1) * RQ;
        if (lane == 0) {
            for (int st = 0; st < NSTG; ++st) mbar_init(bar0 + 8u * st, 1);
            fence_barrier_init();
        }
        for (int i = lane; i < PW * NF; i += 32) zero_s[i] = 0.0;     // (every warp writes the same zeros: no CTA barrier needed)
        __syncwarp();
    }
    auto ring_issue = [&](const int j, const int st) {         // block j into stage st (whole warp calls)
        const uint32_t bar = bar0 + 8u * st;
        if (lane == 0) mbar_expect_tx(bar, SM::SGB);
        __syncwarp();
        if (lane < CPW) {
            const double *src = csrc + (int64_t)j * PW * NF;
            bulk_g2s(smem_u32(wreg + st * SM::SGB + lane * SM::CBR), reinterpret_cast<const double *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)15), SM::CBR, bar);
        } else if (R > 0 && lane < 2 * CPW) {
            const double *src = csrc + (int64_t)j * PW * R;
            bulk_g2s(smem_u32(wreg + st * SM::SGB + CPW * SM::CBR + (lane - CPW) * SM::CBV), reinterpret_cast<const double *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)15), SM::CBV, bar);
        }
    };
    // this lane's view of block j in stage st: first operand row / first V row of its own chunk
    auto ring_rec = [&](const int j, const int st) -> const double * {
        const uintptr_t g = reinterpret_cast<uintptr_t>(rec0 + (int64_t)(LP + j * PW) * NF);
        return reinterpret_cast<const double *>(wreg + st * SM::SGB + grp * SM::CBR + (g & 15));
    };
    auto ring_v = [&](const int j, const int st) -> const double * {
        const uintptr_t g = reinterpret_cast<uintptr_t>(v0 + (int64_t)(LP + 1 + j * PW) * R);
        return reinterpret_cast<const double *>(wreg + st * SM::SGB + CPW * SM::CBR + grp * SM::CBV + (g & 15));
    };
    // rhs values of block j -> b buffer of stage st: lane gl of a group fetches entries gl, gl + GS, ... of its chunk's block
    double *bbuf = reinterpret_cast<double *>(wreg + SM::BOFF) + grp * SM::PWE;
    auto b_issue = [&](const int j, const int st) {
#pragma unroll
        for (int t = gl; t < PW; t += GS) {
            const int64_t i = (int64_t)LP + (int64_t)j * PW + t;           // row of the chunk
            const bool valid = j < nblk_ring && i < brows;
            const double *src = valid ? b0 + i : a.b;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(bbuf + st * CPW * SM::PWE + t)), "l"(src), "r"(valid ? 8 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int rstage = 0, rblock = 0; uint32_t rparity = 0;          // stage / block being read, parity of that stage's barrier
    {
        for (int j = 0; j < NSTG && j < nblk_ring; ++j) ring_issue(j, j);
        for (int j = 0; j < NSTG; ++j) b_issue(j, j);          // always NSTG groups in flight (empty past the last block)
        asm volatile("cp.async.wait_group %0;" ::"n"(NSTG - 1) : "memory");
        mbar_wait(bar0, 0);
        __syncwarp();
    }

    // column roles of this lane's slots
    // COLOC: the R generator columns sit in ONE lane (consecutive slots of lane G0 % GS) instead of R neighbouring lanes, so the
    // fill value of a surviving row is formed where the generator lives and travels in one shuffle instead of R (the LSU pipe
    // that executes shuffles is the busiest unit of this kernel).  Generator column q > 0 trades places with whatever canonical
    // column (or spare position) its new position held.
    constexpr bool COLOC = R >= 2 && (G0 / GS + R <= CS);
    constexpr int GLANE = G0 % GS, GSLOT = G0 / GS;
    bool isW[CS], isG[CS], isC[CS], isT[CS], isB[CS];
    int col[CS];
#pragma unroll
    for (int sl = 0; sl < CS; ++sl) {
        col[sl] = sl * GS + gl;
        if (COLOC) {
#pragma unroll
            for (int q = 1; q < R; ++q) {
                const int moved = (GSLOT + q) * GS + GLANE;       // position that now holds generator column q
                if (sl * GS + gl == moved) col[sl] = G0 + q;
                else if (sl * GS + gl == G0 + q) col[sl] = moved;
            }
        }
        isW[sl] = col[sl] < PW; isG[sl] = col[sl] >= G0 && col[sl] < C0; isC[sl] = col[sl] >= C0 && col[sl] < T0;
        isT[sl] = col[sl] >= T0 && col[sl] < B0; isB[sl] = col[sl] == B0;
    }
    double A[CS][NR];
#pragma unroll
    for (int sl = 0; sl < CS; ++sl)
#pragma unroll
        for (int s = 0; s < NR; ++s) A[sl][s] = 0.0;

    // generator value q of row slot s, broadcast inside the group
    auto gen_of = [&](int s, int q) -> double {
        return COLOC ? __shfl_sync(FULL, A[GSLOT + q][s], GLANE, GS) : __shfl_sync(FULL, A[(G0 + q) / GS][s], (G0 + q) % GS, GS);
    };

    // ---- prologue: band rows 0..LP-1 (slot = row) and the R tail rows (slots PW..) ----
#pragma unroll
    for (int ip = 0; ip < LP; ++ip) {
        const double *rec = rec0 + (int64_t)ip * NF;
#pragma unroll
        for (int sl = 0; sl < CS; ++sl) {
            double val = 0.0;
            if (isW[sl] && col[sl] <= ip) val = rec[LP - ip + col[sl]];
            if (isG[sl]) val = rec[PW + col[sl] - G0];
            if (isC[sl] && ip <= col[sl] - C0) val = rec[col[sl] - C0 - ip];
            if (isB[sl] && ip < brows) val = b0[ip];
            A[sl][ip] = val;
        }
        double gq[RQ];
#pragma unroll
        for (int q = 0; q < R; ++q) gq[q] = gen_of(ip, q);
#pragma unroll
        for (int sl = 0; sl < CS; ++sl) {
            if (sl * GS >= PW) continue;
            double f = 0.0;
#pragma unroll
            for (int q = 0; q < R; ++q) f = fma(gq[q], isW[sl] ? v0[(int64_t)col[sl] * R + q] : 0.0, f);
            if (isW[sl] && col[sl] > ip) A[sl][ip] = f;
        }
    }
#pragma unroll
    for (int t = 0; t < R; ++t)
#pragma unroll
        for (int sl = 0; sl < CS; ++sl) {
            double val = 0.0;
            if (isW[sl]) val = -v0[(int64_t)col[sl] * R + t];
            if (isG[sl] && col[sl] - G0 == t) val = -1.0;
            if (isT[sl] && col[sl] - T0 == t) val = 1.0;
            A[sl][PW + t] = val;
        }

    // Per-slot stream of the entering rows: window / generator columns read the row records of the current block in the ring (their
    // pointer is re-based at every block, the row inside the block is a compile-time offset), the rhs column reads the rhs buffer,
    // all other columns read a block of zero rows in shared memory, so the load needs neither a predicate nor a per-slot stride.
    // A window column reads field d = (col - k) mod PW, which is also the position of its entry in the emitted record.
    // Emitted records: one running pointer for the warp's row, a constant element offset per slot.
    int d[CS]; const double *ep[CS]; double nxt[CS]; int fo[CS]; bool emit[CS];
    double *frb = TILED ? a.FR + (int64_t)(c >> 2) * a.npf * NC * 4 + (c & 3) : a.FR + r0 * NC;
    const double *bcur = bbuf;                                // this block's rhs values
    double bnx = bcur[0];
#pragma unroll
    for (int sl = 0; sl < CS; ++sl) {
        d[sl] = isW[sl] ? col[sl] : 0;                        // phase 0
        if (isW[sl] || isG[sl]) ep[sl] = ring_rec(0, 0) + (isG[sl] ? PW + col[sl] - G0 : 0);
        else ep[sl] = zero_s;
        nxt[sl] = ep[sl][d[sl]];
        fo[sl] = (isW[sl] ? 0 : col[sl]) * FSTR;
        emit[sl] = live && col[sl] < NC;
    }
    const double *vp = ring_v(0, 0);                          // V row of the column entering after step 0
    double vnx[RQ];
#pragma unroll
    for (int q = 0; q < R; ++q) vnx[q] = vp[q];

    for (int kb = 0; kb < nsteps; kb += PW) {
#pragma unroll
        for (int PH = 0; PH < PW; ++PH) {
            const int es = (PH + LP) % PW;                    // slot of the entering row
            const int pl = PH % GS, ps = PH / GS;             // lane-in-group / column slot of the pivot column
#pragma unroll
            for (int sl = 0; sl < CS; ++sl) A[sl][es] = isB[sl] ? bnx : nxt[sl];
            if (PH == PW - 1) {
                // the rows fetched from here on belong to the next block: wait for its stage, hand this block's stage back
                const int old = rstage;
                if (++rstage == NSTG) { rstage = 0; rparity ^= 1u; }
                ++rblock;
                asm volatile("cp.async.wait_group %0;" ::"n"(NSTG - 2) : "memory");    // the next block's rhs values have landed
                mbar_wait(bar0 + 8u * rstage, rparity);
                __syncwarp();                                 // every lane's reads of the old stage are in registers ...
                fence_proxy_async();                          // ... and ordered before the bulk copies that re-arm it
                if (rblock - 1 + NSTG < nblk_ring) ring_issue(rblock - 1 + NSTG, old);
                b_issue(rblock - 1 + NSTG, old);
                bcur = bbuf + rstage * CPW * SM::PWE;
            }
            bnx = bcur[(PH + 1) % PW];
#pragma unroll
            for (int sl = 0; sl < CS; ++sl) {
                // prefetch the row entering at the next step (rows past the chunk are the next chunk's or the identity pad)
                if (PH == PW - 1) { if (isW[sl] || isG[sl]) ep[sl] = ring_rec(rblock, rstage) + (isG[sl] ? PW + col[sl] - G0 : 0); }
                const int dn = isW[sl] ? (d[sl] == 0 ? LP : d[sl] - 1) : 0;
                nxt[sl] = ep[sl][((PH + 1) % PW) * NF + dn];
            }
            double vn[RQ];
            if (PH == PW - 1) vp = ring_v(rblock, rstage); else vp += R;
#pragma unroll
            for (int q = 0; q < R; ++q) { vn[q] = vnx[q]; vnx[q] = vp[q]; }   // one step ahead
            // reflector from the pivot column, LinearAlgebra.reflector! convention
            double x[NR];
#pragma unroll
            for (int s = 0; s < NR; ++s) x[s] = __shfl_sync(FULL, A[ps][s], pl, GS);
            double q0 = 0.0, q1 = 0.0, q2 = 0.0;
#pragma unroll
            for (int s = 0; s < NR; ++s) { if (s % 3 == 0) q0 = fma(x[s], x[s], q0); else if (s % 3 == 1) q1 = fma(x[s], x[s], q1); else q2 = fma(x[s], x[s], q2); }
            // projections of this lane's columns on the RAW pivot column: they need nothing of the reflector's scalars, so they run
            // beside the rsqrt / rcp chain instead of behind it (v_s = x_s * inv is never formed; inv joins at the end)
            double raw[CS];
#pragma unroll
            for (int sl = 0; sl < CS; ++sl) {
                double d0 = 0.0, d1 = 0.0;
#pragma unroll
                for (int s = 0; s < NR; ++s) {
                    if (s == PH) continue;
                    if (s & 1) d1 = fma(x[s], A[sl][s], d1); else d0 = fma(x[s], A[sl][s], d0);
                }
                raw[sl] = d0 + d1;
            }
            const double ssq = (q0 + q1) + q2;
            const double xp = x[PH];
            const bool nz = ssq != 0.0;
            const double rs = fast_rsqrt(ssq);
            double nrm = ssq * rs;
            nrm = fma(fma(-nrm, nrm, ssq), 0.5 * rs, nrm);    // one correction step: sqrt to ~1 ulp
            const double nu = copysign(nrm, xp);
            const double xi = xp + nu;
            double inv, tau;
            if constexpr (BF) { inv = keep_if(fast_rcp(xi), nz); tau = keep_if(xi * copysign(rs, xp), nz); }
            else { inv = nz ? fast_rcp(xi) : 0.0; tau = nz ? xi * copysign(rs, xp) : 0.0; }
#pragma unroll
            for (int sl = 0; sl < CS; ++sl) {
                const double dot = fma(inv, raw[sl], A[sl][PH]) * tau;     // tau * v'a, v = (1, x_s * inv)
                const double cf = dot * inv;
#pragma unroll
                for (int s = 0; s < NR; ++s) A[sl][s] = (s == PH) ? A[sl][s] - dot : fma(-x[s], cf, A[sl][s]);
            }
            if (gl == pl) A[ps][PH] = nz ? -nu : xp;
            // emit the finished row (canonical order: window entry of column k+t at position t)
#pragma unroll
            for (int sl = 0; sl < CS; ++sl) {
                double *dst = frb + (fo[sl] + (isW[sl] ? d[sl] : 0) * FSTR);
                if constexpr (BF) st_global_if(dst, A[sl][PH], emit[sl]);
                else if (emit[sl]) *dst = A[sl][PH];
                d[sl] = isW[sl] ? (d[sl] == 0 ? LP : d[sl] - 1) : 0;
            }
            frb += RSTR;
            // the window slides: local column k+LP+1 replaces the pivot column (fill region of every surviving row)
#pragma unroll
            for (int s = 0; s < NR; ++s) {
                if (s == PH) continue;
                double f = 0.0;
                if (COLOC) {
#pragma unroll
                    for (int q = 0; q < R; ++q) f = fma(A[GSLOT + q][s], vn[q], f);     // meaningful in the generator lane only
                    f = __shfl_sync(FULL, f, GLANE, GS);
                } else {
#pragma unroll
                    for (int q = 0; q < R; ++q) f = fma(gen_of(s, q), vn[q], f);
                }
                if (gl == pl) A[ps][s] = f;
            }
        }
    }
    // ---- leftover rows -> level-0 reduced block  [F = C_L | Tprev,  G = window(S_{c+1}) | generator,  h] ----
#pragma unroll
    for (int sl = 0; sl < CS; ++sl)
#pragma unroll
        for (int s = 0; s < NR; ++s) stage[sl][s][lane] = A[sl][s];       // (the ring has drained: every issued block was waited for)
    __syncwarp();
    if (!live) return;
    double *E = a.E0 + (int64_t)c * W * (2 * W + 1);
#pragma unroll
    for (int sl = 0; sl < CS; ++sl) {
        int dst = -1;
        if (isW[sl]) {                             // window position col holds local column jp == col (mod PW) in [m-LP, m]; column m is not part of the chunk
            int jp = (m - LP) + ((col[sl] - (m - LP)) % PW + PW) % PW;
            if (jp <= m - 1) dst = W + (jp - (m - LP));
        } else if (isG[sl]) dst = W + LP + (col[sl] - G0);
        else if (isC[sl]) dst = col[sl] - C0;
        else if (isT[sl]) dst = LP + (col[sl] - T0);
        else if (isB[sl]) dst = 2 * W;
        if (dst >= 0) {
            for (int row = 0; row < W; ++row) {
                const int slot = row < LP ? (m - LP + row) % PW : PW + (row - LP);
                E[row * (2 * W + 1) + dst] = stage[sl][slot][lane];
            }
        }
    }
}

// Chunk-interleaved copy of the rows and V rows that enter the sweep of stage 1, for the lane-per-chunk kernel: tile t = chunks
// 32t .. 32t+31 (equally long), step i = 0 .. nsteps-1:  RI[((t * nrow + i) * NF + f) * 32 + lane] = REC[r0 + LP + i][f],
// VI[((t * nrow + i) * R + q) * 32 + lane] = V[r0 + u + LP + 1 + i][q]  (r0 = first row of chunk 32t + lane; spare lanes repeat chunk P-1).
template <int NF, int R>
__global__ void __launch_bounds__(256) abc_interleave_kernel(const AbcArgs a, double *RI, double *VI, const int lp, const int nrow) {
    const int lane = threadIdx.x & 31;
    const int64_t t = blockIdx.y;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= nrow) return;
    int c = (int)(t * 32 + lane);
    if (c >= a.P) c = a.P - 1;
    const int64_t r0 = chunk_start(a, c);
    const double *src = a.REC + (r0 + lp + i) * NF, *vs = a.V + (r0 + a.u + lp + 1 + i) * R;
    double *dst = RI + ((t * nrow + i) * NF) * 32 + lane, *vd = VI + ((t * nrow + i) * R) * 32 + lane;
#pragma unroll
    for (int f = 0; f < NF; ++f) dst[f * 32] = src[f];
#pragma unroll
    for (int q = 0; q < R; ++q) vd[q * 32] = vs[q];
}

// ---------------------------------------------------------------------------------------------------------------
// stage 1, one LANE per chunk (shapes (l+u, r) = (8 | 6 | 4, 2) at many chunks).  A CTA is FOUR warps sweeping the same 32 chunks,
// lane = chunk in every warp, so there is no pivot broadcast, no redundant reflector arithmetic and no shuffle at all; the 22
// columns of a chunk's active block are spread over the four warps' registers:
//   P1  window positions 0..NP1-1   forms the reflector of every step (the pivot column is always position 0) and publishes
//                                   (x, 1/xi, tau) plus the entering row's remaining fields and the entering V row
//   P2  window positions NP1+1..l+u | generator columns     applies it, builds the column that enters the window, and hands the column
//                                   that moves from position NP1+1 to NP1 over to P1 (one column per step, through shared memory)
//   T1, T2   separator | tail | rhs columns                  apply it
// The window is kept in RELATIVE position — column k+j of step k at position j, row k+i in slot i — and every update writes its
// result one position / one slot down (a different destination register, no extra instruction), so the step is the same code for
// every k and the loop is not unrolled over k mod (l+u+1).  Rings: reflectors P1 -> {P2, T1, T2} (NSX slots, full / empty
// mbarriers, one arrival per warp), moving column P2 -> P1 (2 slots), entering operand and V rows by TWO bulk-async copies per
// stage for the whole warp — they come from a chunk-interleaved copy of the operands, RI[tile of 32 chunks][step][field][lane]
// and VI[tile][step][r][lane], that abc_interleave_kernel writes once per operand set and chunk layout (feeding 32 separate
// row-major chunk streams into one warp costs a bulk copy per lane and stage: measured, profiles/r2_measurements.md); rhs values
// by plain loads two steps ahead (T2).  Same arithmetic, R records and level-0 blocks as abc_qr_kernel.
// ---------------------------------------------------------------------------------------------------------------
template <int LP, int R>
struct AbcLpcSmem {
    using D = AbcDims<LP, R>;
    static constexpr int NP1 = (D::PW + 1) / 2;                     // window positions 0..NP1-1 live in P1 (5 of 9 at l+u = 8); position NP1 arrives every step
    static constexpr int NT1 = (LP + R + 1 + 1) / 2;                // trailing columns of T1 (the rest, with the rhs, in T2)
    static constexpr int SBK = D::PW % 3 == 0 ? 3 : D::PW;          // steps per stage of the entering-row ring (a divisor of the chunks' step counts)
    static_assert(D::PW % SBK == 0 && NP1 + 1 < D::PW, "ring stage / column split");
    static constexpr int NSTG = 2, NSX = 4, NSC = 2;
    static constexpr int CBR = SBK * D::NF * 256, CBV = SBK * R * 256;           // a stage: SBK rows of the tile-interleaved operand / V rows
    static constexpr int SGB = CBR + CBV;
    static constexpr int XF = (D::NR - 1) + 2;                      // x[1 .. NR-1], 1/xi, tau
    static constexpr int XN = XF + (D::NF - NP1 - 1) + R;           // + entering row fields NP1+1 .. NF-1, entering V row
    static constexpr int CN = D::NR - 1;                            // moving column: every slot but the entering row's
    static constexpr int OFF_RING = 0, OFF_X = OFF_RING + NSTG * SGB, OFF_C = OFF_X + NSX * XN * 256, BARS = OFF_C + NSC * CN * 256;
    static constexpr int NBAR = NSTG + 2 * NSX + 2 * NSC;           // entering-row ring | X full | X empty | C full | C empty
    static constexpr int TOTAL = BARS + NBAR * 8;
};

template <int LP, int R, bool TILED>
__global__ void __launch_bounds__(128, 3) abc_qr_lpc_kernel(const AbcArgs a) {
    using D = AbcDims<LP, R>;
    using SM = AbcLpcSmem<LP, R>;
    static_assert(R >= 1, "built for shapes with a fill");
    constexpr int PW = D::PW, NR = D::NR, NC = D::NC, NF = D::NF, W = D::W, C0 = D::C0;
    constexpr int NP1 = SM::NP1, NP2 = PW - NP1 - 1, NCB = LP + R + 1, NT1 = SM::NT1, NT2 = NCB - NT1;
    constexpr int SBK = SM::SBK, NSTG = SM::NSTG, NSX = SM::NSX, NSC = SM::NSC, XF = SM::XF, XN = SM::XN, CN = SM::CN;
    static_assert(TILED, "the lane-per-chunk stage 1 writes records interleaved over its tile of 32 chunks");
    constexpr int FSTR = 32, RSTR = NC * 32;                  // FR[tile][k][field][lane]: a field of one step is one coalesced 256-byte store
    extern __shared__ __align__(128) unsigned char smem[];
    // The four roles rotate over the CTA's warps with the order in which CTAs arrive on this SM (a counter per SM), so that the CTAs
    // resident together give every scheduler a mix of roles wherever the block scheduler puts them
    __shared__ int rot_s;
    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        rot_s = atomicAdd(a.smc + (smid & 511), 1) & 3;
    }
    if (threadIdx.x < SM::NBAR) {
        const int i = threadIdx.x;
        mbar_init(smem_u32(smem + SM::BARS + i * 8), (i >= NSTG + NSX && i < NSTG + 2 * NSX) ? 3 : 1);   // X empty: three consumer warps
    }
    fence_barrier_init();
    __syncthreads();
    const int lane = threadIdx.x & 31, role = ((threadIdx.x >> 5) + rot_s) & 3;     // 0 P1, 1 P2, 2 T1, 3 T2
    int c = blockIdx.x * 32 + lane;
    const bool live = c < a.P;
    if (!live) c = a.P - 1;
    const int64_t r0 = chunk_start(a, c);
    const int m = a.m0 + (c < a.jx ? PW : 0), nsteps = m - LP, nsb = nsteps / SBK;   // the same for the 32 chunks of a CTA
    const int u = a.u;
    const double *rec0 = a.REC + r0 * NF;
    const double *v0 = a.V + (r0 + u) * R;
    const double *b0 = a.b + r0;
    const int64_t brows = a.n - r0;
    const uint32_t bar_ring = smem_u32(smem + SM::BARS), bar_xf = bar_ring + NSTG * 8, bar_xe = bar_xf + NSX * 8, bar_cf = bar_xe + NSX * 8, bar_ce = bar_cf + NSC * 8;
    double *frb = a.FR + (int64_t)blockIdx.x * a.npf * NC * 32 + lane;
    double *E = a.E0 + (int64_t)c * W * (2 * W + 1);
    double *Xs = reinterpret_cast<double *>(smem + SM::OFF_X) + lane;           // Xs[(slot * XN + f) * 32]
    double *Cs = reinterpret_cast<double *>(smem + SM::OFF_C) + lane;           // Cs[(slot * CN + i) * 32], i = slot index without the entering row's
    // value of window column j (local), band row ip < LP, before any elimination: band entry, or fill U[ip,:] . V[:, j]
    auto wval = [&](const int ip, const int j) -> double {
        const double *rec = rec0 + (int64_t)ip * NF;
        if (j <= ip) return rec[LP - ip + j];
        double val = 0.0;
#pragma unroll
        for (int q = 0; q < R; ++q) val = fma(rec[PW + q], v0[(int64_t)j * R + q], val);
        return val;
    };
    // one column through one reflector (x[1..], inv, tau): cv has the entering entry in slot LP; band rows move one slot down
#define ABC_LPC_REFLECT(cv, out, ev)                                                                   \
    {                                                                                                  \
        double d0_ = 0.0, d1_ = 0.0;                                                                   \
        _Pragma("unroll") for (int s = 1; s < NR; ++s) { if (s & 1) d1_ = fma(x[s], (cv)[s], d1_); else d0_ = fma(x[s], (cv)[s], d0_); } \
        const double dot_ = fma(inv, d0_ + d1_, (cv)[0]) * tau;                                        \
        const double cf_ = dot_ * inv;                                                                 \
        ev = (cv)[0] - dot_;                                                                           \
        _Pragma("unroll") for (int s = 1; s < PW; ++s) (out)[s - 1] = fma(-x[s], cf_, (cv)[s]);        \
        _Pragma("unroll") for (int s = PW; s < NR; ++s) (out)[s] = fma(-x[s], cf_, (cv)[s]);           \
    }

    if (role == 0) {
        // ===== P1: window positions 0..NP1-1, the reflector =====
        unsigned char *ring = smem + SM::OFF_RING;
        const double *ri_t = a.RI + (int64_t)blockIdx.x * a.nrow_i * NF * 32, *vi_t = a.VI + (int64_t)blockIdx.x * a.nrow_i * R * 32;
        auto ring_issue = [&](const int j, const int st) {
            if (lane == 0) {
                const uint32_t bar = bar_ring + 8u * st;
                mbar_expect_tx(bar, SM::SGB);
                bulk_g2s(smem_u32(ring + st * SM::SGB), ri_t + (int64_t)j * SBK * NF * 32, SM::CBR, bar);
                bulk_g2s(smem_u32(ring + st * SM::SGB + SM::CBR), vi_t + (int64_t)j * SBK * R * 32, SM::CBV, bar);
            }
        };
        for (int j = 0; j < NSTG && j < nsb; ++j) ring_issue(j, j);
        double Wr[NP1][NR];
#pragma unroll
        for (int j = 0; j < NP1; ++j) {
#pragma unroll
            for (int ip = 0; ip < LP; ++ip) Wr[j][ip] = wval(ip, j);
            Wr[j][LP] = 0.0;
#pragma unroll
            for (int t = 0; t < R; ++t) Wr[j][PW + t] = -v0[(int64_t)j * R + t];
        }
        int st = 0, xs = 0, cs = 0; uint32_t par = 0, xpar = 1, cpar = 0;
        for (int jb = 0; jb < nsb; ++jb) {
            mbar_wait(bar_ring + 8u * st, par);
            const double *rb = reinterpret_cast<const double *>(ring + st * SM::SGB) + lane;              // rb[(i * NF + f) * 32]
            const double *vb = reinterpret_cast<const double *>(ring + st * SM::SGB + SM::CBR) + lane;    // vb[(i * R + q) * 32]
#pragma unroll
            for (int i = 0; i < SBK; ++i) {
                auto rec = [&](const int f) -> double { return rb[(i * NF + f) * 32]; };       // entering row, read where it is used
                double x[NR];
#pragma unroll
                for (int s = 0; s < NR; ++s) x[s] = Wr[0][s];
                x[LP] = rec(0);                               // the entering row (local row k + LP) sits in slot LP
                double q0 = 0.0, q1 = 0.0, q2 = 0.0;
#pragma unroll
                for (int s = 0; s < NR; ++s) { if (s % 3 == 0) q0 = fma(x[s], x[s], q0); else if (s % 3 == 1) q1 = fma(x[s], x[s], q1); else q2 = fma(x[s], x[s], q2); }
                const double ssq = (q0 + q1) + q2;
                const double xp = x[0];
                const bool nz = ssq != 0.0;
                const double rs = fast_rsqrt(ssq);
                double nrm = ssq * rs;
                nrm = fma(fma(-nrm, nrm, ssq), 0.5 * rs, nrm);
                const double nu = copysign(nrm, xp);
                const double xi = xp + nu;
                const double inv = nz ? fast_rcp(xi) : 0.0;
                const double tau = nz ? xi * copysign(rs, xp) : 0.0;
                // publish the reflector, the rest of the entering row and the entering V row
                mbar_wait(bar_xe + 8u * xs, xpar);
                {
                    double *xo = Xs + xs * XN * 32;
#pragma unroll
                    for (int s = 1; s < NR; ++s) xo[(s - 1) * 32] = x[s];
                    xo[(NR - 1) * 32] = inv; xo[NR * 32] = tau;
#pragma unroll
                    for (int f = NP1 + 1; f < NF; ++f) xo[(XF + f - NP1 - 1) * 32] = rec(f);
#pragma unroll
                    for (int q = 0; q < R; ++q) xo[(XF + NF - NP1 - 1 + q) * 32] = vb[(i * R + q) * 32];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_xf + 8u * xs);
                if (++xs == NSX) { xs = 0; xpar ^= 1u; }
                if (live) frb[0] = nz ? -nu : xp;
#pragma unroll
                for (int j = 1; j < NP1; ++j) {
                    Wr[j][LP] = rec(j);
                    double ev;
                    ABC_LPC_REFLECT(Wr[j], Wr[j - 1], ev)
                    if (live) frb[j * FSTR] = ev;
                }
                {   // the column at position NP1: handed over by P2 after the previous step (by its prologue for step 0)
                    double cv[NR], ev;
                    mbar_wait(bar_cf + 8u * cs, cpar);
                    const double *ci = Cs + cs * CN * 32;
#pragma unroll
                    for (int s = 0; s < NR; ++s) cv[s] = s == LP ? rec(NP1) : ci[(s < LP ? s : s - 1) * 32];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_ce + 8u * cs);
                    if (++cs == NSC) { cs = 0; cpar ^= 1u; }
                    ABC_LPC_REFLECT(cv, Wr[NP1 - 1], ev)
                    if (live) frb[NP1 * FSTR] = ev;
                }
                frb += RSTR;
            }
            __syncwarp();
            fence_proxy_async();
            if (jb + NSTG < nsb) ring_issue(jb + NSTG, st);
            if (++st == NSTG) { st = 0; par ^= 1u; }
        }
        // leftover rows -> window part of the level-0 block, positions 0..NP1 (the last one arrives from P2 once more)
        double cl[NR];
        mbar_wait(bar_cf + 8u * cs, cpar);
        {
            const double *ci = Cs + cs * CN * 32;
#pragma unroll
            for (int s = 0; s < NR; ++s) cl[s] = s == LP ? 0.0 : ci[(s < LP ? s : s - 1) * 32];
        }
        if (!live) return;
#pragma unroll
        for (int row = 0; row < W; ++row) {
            const int slot = row < LP ? row : PW + (row - LP);
#pragma unroll
            for (int j = 0; j < NP1; ++j) E[row * (2 * W + 1) + W + j] = Wr[j][slot];
            E[row * (2 * W + 1) + W + NP1] = cl[slot];
        }
    } else if (role == 1) {
        // ===== P2: window positions NP1+1..LP (registers Wp[0..NP2-1]) | generator columns =====
        double Wp[NP2][NR], G[R][NR];
#pragma unroll
        for (int ip = 0; ip < LP; ++ip) {
#pragma unroll
            for (int q = 0; q < R; ++q) G[q][ip] = rec0[(int64_t)ip * NF + PW + q];
#pragma unroll
            for (int j = 0; j < NP2; ++j) Wp[j][ip] = wval(ip, NP1 + 1 + j);
        }
#pragma unroll
        for (int t = 0; t < R; ++t) {
#pragma unroll
            for (int j = 0; j < NP2; ++j) Wp[j][PW + t] = -v0[(int64_t)(NP1 + 1 + j) * R + t];
#pragma unroll
            for (int q = 0; q < R; ++q) G[q][PW + t] = q == t ? -1.0 : 0.0;
        }
#pragma unroll
        for (int j = 0; j < NP2; ++j) Wp[j][LP] = 0.0;
#pragma unroll
        for (int q = 0; q < R; ++q) G[q][LP] = 0.0;
        int xs = 0, cs = 0; uint32_t xpar = 0, cpar = 1;
        {   // message 0 of the column ring: position NP1 as the prologue leaves it
            mbar_wait(bar_ce + 8u * cs, cpar);
            double *co = Cs + cs * CN * 32;
#pragma unroll
            for (int ip = 0; ip < LP; ++ip) co[ip * 32] = wval(ip, NP1);
#pragma unroll
            for (int t = 0; t < R; ++t) co[(LP + t) * 32] = -v0[(int64_t)NP1 * R + t];
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_cf + 8u * cs);
            if (++cs == NSC) { cs = 0; cpar ^= 1u; }
        }
        for (int k = 0; k < nsteps; ++k) {
            double x[NR], inv, tau, rec[NF], vr[R];
            mbar_wait(bar_xf + 8u * xs, xpar);
            {
                const double *xi_ = Xs + xs * XN * 32;
#pragma unroll
                for (int s = 1; s < NR; ++s) x[s] = xi_[(s - 1) * 32];
                inv = xi_[(NR - 1) * 32]; tau = xi_[NR * 32];
#pragma unroll
                for (int f = NP1 + 1; f < NF; ++f) rec[f] = xi_[(XF + f - NP1 - 1) * 32];
#pragma unroll
                for (int q = 0; q < R; ++q) vr[q] = xi_[(XF + NF - NP1 - 1 + q) * 32];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_xe + 8u * xs);
            if (++xs == NSX) { xs = 0; xpar ^= 1u; }
            {   // position NP1+1 moves to NP1: over to P1
                double out[NR], ev;
                Wp[0][LP] = rec[NP1 + 1];
                ABC_LPC_REFLECT(Wp[0], out, ev)
                if (live) frb[(NP1 + 1) * FSTR] = ev;
                mbar_wait(bar_ce + 8u * cs, cpar);
                double *co = Cs + cs * CN * 32;
#pragma unroll
                for (int s = 0; s < NR; ++s) if (s != LP) co[(s < LP ? s : s - 1) * 32] = out[s];
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_cf + 8u * cs);
                if (++cs == NSC) { cs = 0; cpar ^= 1u; }
            }
#pragma unroll
            for (int j = 1; j < NP2; ++j) {
                Wp[j][LP] = rec[NP1 + 1 + j];
                double ev;
                ABC_LPC_REFLECT(Wp[j], Wp[j - 1], ev)
                if (live) frb[(NP1 + 1 + j) * FSTR] = ev;
            }
#pragma unroll
            for (int q = 0; q < R; ++q) {
                G[q][LP] = rec[PW + q];
                double ev;
                ABC_LPC_REFLECT(G[q], G[q], ev)
                if (live) frb[(PW + q) * FSTR] = ev;
            }
            // the column that enters the window at position LP: fill region of every surviving row
#pragma unroll
            for (int s = 0; s < NR; ++s) {
                if (s == LP) continue;
                double f = 0.0;
#pragma unroll
                for (int q = 0; q < R; ++q) f = fma(G[q][s], vr[q], f);
                Wp[NP2 - 1][s] = f;
            }
            frb += RSTR;
        }
        if (!live) return;
        // leftover rows -> window positions NP1+1..LP-1 and generator part of the level-0 block
#pragma unroll
        for (int row = 0; row < W; ++row) {
            const int slot = row < LP ? row : PW + (row - LP);
#pragma unroll
            for (int j = 0; j + NP1 + 1 < LP; ++j) E[row * (2 * W + 1) + W + NP1 + 1 + j] = Wp[j][slot];
#pragma unroll
            for (int q = 0; q < R; ++q) E[row * (2 * W + 1) + W + LP + q] = G[q][slot];
        }
    } else {
        // ===== T1 / T2: trailing columns jt (separator jt < LP, tail LP <= jt < LP + R, rhs jt = LP + R) =====
        // (one body per warp, with its column range as compile-time constants: a run-time column count would either cost T2 a reflector
        // of zeros per step or put a branch into the column loop, and the branch costs the overlap of the reflectors — measured)
        auto trailing = [&](auto jt0_tag, auto cnt_tag) {
            constexpr int JT0 = decltype(jt0_tag)::value, NTC = decltype(cnt_tag)::value;
            constexpr bool HASRHS = JT0 + NTC == NCB;
            double Tr[NTC][NR];
#pragma unroll
            for (int j = 0; j < NTC; ++j) {
                const int jt = JT0 + j;
#pragma unroll
                for (int s = 0; s < NR; ++s) {
                    double val = 0.0;
                    if (s < LP) {
                        if (jt < LP) { if (s <= jt) val = rec0[(int64_t)s * NF + (jt - s)]; }
                        else if (jt == LP + R) { if (s < brows) val = b0[s]; }
                    } else if (s >= PW) {
                        if (jt >= LP && jt < LP + R && jt - LP == s - PW) val = 1.0;
                    }
                    Tr[j][s] = val;
                }
            }
            auto bload = [&](const int64_t i) -> double { return (HASRHS && i < brows) ? b0[i] : 0.0; };
            double bq0 = bload(LP), bq1 = bload(LP + 1), bq2 = bload(LP + 2);
            int xs = 0; uint32_t xpar = 0;
            for (int k = 0; k < nsteps; ++k) {
                double x[NR], inv, tau;
                mbar_wait(bar_xf + 8u * xs, xpar);
                {
                    const double *xi_ = Xs + xs * XN * 32;
#pragma unroll
                    for (int s = 1; s < NR; ++s) x[s] = xi_[(s - 1) * 32];
                    inv = xi_[(NR - 1) * 32]; tau = xi_[NR * 32];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_xe + 8u * xs);
                if (++xs == NSX) { xs = 0; xpar ^= 1u; }
                const double bval = bq0;
                bq0 = bq1; bq1 = bq2; bq2 = bload((int64_t)k + LP + 3);
#pragma unroll
                for (int j = 0; j < NTC; ++j) {
                    Tr[j][LP] = (JT0 + j == LP + R) ? bval : 0.0;
                    double ev;
                    ABC_LPC_REFLECT(Tr[j], Tr[j], ev)
                    if (live) frb[(C0 + JT0 + j) * FSTR] = ev;
                }
                frb += RSTR;
            }
            if (!live) return;
#pragma unroll
            for (int row = 0; row < W; ++row) {
                const int slot = row < LP ? row : PW + (row - LP);
#pragma unroll
                for (int j = 0; j < NTC; ++j) {
                    const int jt = JT0 + j;
                    E[row * (2 * W + 1) + (jt < W ? jt : 2 * W)] = Tr[j][slot];
                }
            }
        };
        if (role == 2) trailing(std::integral_constant<int, 0>{}, std::integral_constant<int, NT1>{});
        else trailing(std::integral_constant<int, NT1>{}, std::integral_constant<int, NT2>{});
    }
#undef ABC_LPC_REFLECT
}

// ---------------------------------------------------------------------------------------------------------------
// stage 2: merge two adjacent blocks, eliminating the separator between them
//   Ein  : level blocks [W][2W+1] = [F | G | h];  merge i combines blocks 2i, 2i+1
//   TOP  : [W][3W+1] = [Rhat | Fhat | Ghat | hhat]  (z_M = Rhat^-1 (hhat - Fhat z_L - Ghat z_R))
// ---------------------------------------------------------------------------------------------------------------
// Reflector of the column x[j..N-1] (LinearAlgebra.reflector! convention: nu = copysign(||x||, x_j), v = (1, x_s / (x_j + nu)),
// tau = (x_j + nu) / nu) with the branch-free rsqrt / rcp of fastmath.cuh and three-way split sums, as in stage 1; returns the
// projection sum_{s>j} x_s m_s of this lane's column on the RAW pivot column (it does not wait for the scalars).
template <int N>
__device__ __forceinline__ double abc_reflector(const double (&x)[N], const double (&m)[N], const int j, double &inv, double &tau, double &nu, bool &nz) {
    double q0 = 0.0, q1 = 0.0, q2 = 0.0, d0 = 0.0, d1 = 0.0, d2 = 0.0;
#pragma unroll
    for (int s = 0; s < N; ++s) {
        if (s < j) continue;
        if (s % 3 == 0) q0 = fma(x[s], x[s], q0); else if (s % 3 == 1) q1 = fma(x[s], x[s], q1); else q2 = fma(x[s], x[s], q2);
        if (s == j) continue;
        if (s % 3 == 0) d0 = fma(x[s], m[s], d0); else if (s % 3 == 1) d1 = fma(x[s], m[s], d1); else d2 = fma(x[s], m[s], d2);
    }
    const double ssq = (q0 + q1) + q2;
    nz = ssq != 0.0;
    const double rs = fast_rsqrt(ssq);
    double nrm = ssq * rs;
    nrm = fma(fma(-nrm, nrm, ssq), 0.5 * rs, nrm);            // one correction step: sqrt to ~1 ulp
    nu = copysign(nrm, x[j]);
    const double xi = x[j] + nu;
    inv = nz ? fast_rcp(xi) : 0.0;
    tau = nz ? xi * copysign(rs, x[j]) : 0.0;
    return (d0 + d1) + d2;
}

template <int W>
__device__ __forceinline__ void abc_merge_one(const double *Ein, double *Eout, double *TOP, const int nin, const int i, const int lane) {
    constexpr int EW = 2 * W + 1, TW = 3 * W + 1;
    const double *Ea = Ein + (int64_t)(2 * i) * W * EW;
    if (2 * i + 1 >= nin) {                      // odd block passes through
        for (int t = lane; t < W * EW; t += 32) Eout[(int64_t)i * W * EW + t] = Ea[t];
        return;
    }
    const double *Eb = Ea + W * EW;
    double M[2 * W];
#pragma unroll
    for (int s = 0; s < W; ++s) {
        double top = 0.0, bot = 0.0;
        if (lane < W) { top = Ea[s * EW + W + lane]; bot = Eb[s * EW + lane]; }               // z_M: G_a over F_b
        else if (lane < 2 * W) top = Ea[s * EW + (lane - W)];                                  // z_L: F_a
        else if (lane < 3 * W) bot = Eb[s * EW + W + (lane - 2 * W)];                          // z_R: G_b
        else if (lane == 3 * W) { top = Ea[s * EW + 2 * W]; bot = Eb[s * EW + 2 * W]; }
        M[s] = top; M[W + s] = bot;
    }
#pragma unroll
    for (int j = 0; j < W; ++j) {
        double x[2 * W];
#pragma unroll
        for (int s = j; s < 2 * W; ++s) x[s] = __shfl_sync(FULL, M[s], j);
        double inv, tau, nu; bool nz;
        const double xp = x[j];
        const double raw = abc_reflector<2 * W>(x, M, j, inv, tau, nu, nz);
        const double dot = fma(inv, raw, M[j]) * tau;          // tau * v'm, v = (1, x_s * inv)
        const double cf = dot * inv;
        M[j] -= dot;
#pragma unroll
        for (int s = j + 1; s < 2 * W; ++s) M[s] = fma(-x[s], cf, M[s]);
        if (lane == j) {
            M[j] = nz ? -nu : xp;
#pragma unroll
            for (int s = j + 1; s < 2 * W; ++s) M[s] = 0.0;
        }
    }
    if (lane < TW) {
        double *T = TOP + (int64_t)i * W * TW;
#pragma unroll
        for (int s = 0; s < W; ++s) T[s * TW + lane] = M[s];
    }
    if (lane >= W && lane < TW) {
        double *Eo = Eout + (int64_t)i * W * EW;
#pragma unroll
        for (int s = 0; s < W; ++s) Eo[s * EW + (lane - W)] = M[W + s];
    }
}

template <int W>
__global__ void __launch_bounds__(128) abc_merge_kernel(const double *Ein, double *Eout, double *TOP, int nin) {
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= (nin + 1) / 2) return;
    abc_merge_one<W>(Ein, Eout, TOP, nin, i, threadIdx.x & 31);
}

// root: one block [F | G | h] over (z_0, z_P).  Real unknowns: z_0 entries l..W-1 (separator columns >= 0 and the
// free tail T_{-1}) and z_P entries 0..l-1 (separator columns < n); everything else is a phantom and stays 0.
template <int W>
__device__ __forceinline__ void abc_root_one(const double *E, double *Z, const int P, const int l, const int lane, double (*Rm)[W + 1]) {
    constexpr int EW = 2 * W + 1;
    const int n0 = W - l;                         // unknowns taken from z_0
    double M[W];
#pragma unroll
    for (int s = 0; s < W; ++s) {
        double v = 0.0;
        if (lane < n0) v = E[s * EW + l + lane];
        else if (lane < W) v = E[s * EW + W + (lane - n0)];
        else if (lane == W) v = E[s * EW + 2 * W];
        M[s] = v;
    }
#pragma unroll
    for (int j = 0; j < W; ++j) {
        double x[W];
#pragma unroll
        for (int s = j; s < W; ++s) x[s] = __shfl_sync(FULL, M[s], j);
        double inv, tau, nu; bool nz;
        const double xp = x[j];
        const double raw = abc_reflector<W>(x, M, j, inv, tau, nu, nz);
        const double dot = fma(inv, raw, M[j]) * tau;
        const double cf = dot * inv;
        M[j] -= dot;
#pragma unroll
        for (int s = j + 1; s < W; ++s) M[s] = fma(-x[s], cf, M[s]);
        if (lane == j) {
            M[j] = nz ? -nu : xp;
#pragma unroll
            for (int s = j + 1; s < W; ++s) M[s] = 0.0;
        }
    }
    if (lane <= W) {
#pragma unroll
        for (int s = 0; s < W; ++s) Rm[s][lane] = M[s];
    }
    __syncwarp();
    if (lane == 0) {
        double sol[W];
        for (int j = W - 1; j >= 0; --j) {
            double acc = Rm[j][W];
            for (int t = j + 1; t < W; ++t) acc = fma(-Rm[j][t], sol[t], acc);
            sol[j] = acc / Rm[j][j];
        }
        for (int t = 0; t < W; ++t) { Z[t] = 0.0; Z[(int64_t)P * W + t] = 0.0; }
        for (int t = 0; t < n0; ++t) Z[l + t] = sol[t];
        for (int t = n0; t < W; ++t) Z[(int64_t)P * W + (t - n0)] = sol[t];
    }
}

template <int W>
__global__ void __launch_bounds__(32) abc_root_kernel(const double *E, double *Z, int P, int l) {
    __shared__ double Rm[W][W + 1];
    abc_root_one<W>(E, Z, P, l, threadIdx.x, Rm);
}

// expansion of one level: merge i of a level whose blocks span `span` chunks recovers z at chunk index (2i+1)*span.
// One warp per merge: lanes hold the multipliers [z_M (filled in as it is solved) | z_L | z_R | -1] of the stored top rows
// [Rhat | Fhat | Ghat | hhat]; each row is one coalesced load and one butterfly sum.
template <int W>
__device__ __forceinline__ void abc_expand_one(const double *TOP, double *Z, const int i, const int span, const int P, const int lane) {
    constexpr int TW = 3 * W + 1;
    const double *T = TOP + (int64_t)i * W * TW;
    const int64_t iL = (int64_t)2 * i * span, iM = iL + span;
    int64_t iR = iM + span; if (iR > P) iR = P;
    double mult = 0.0;
    if (lane >= W && lane < 2 * W) mult = Z[iL * W + (lane - W)];
    else if (lane >= 2 * W && lane < 3 * W) mult = Z[iR * W + (lane - 2 * W)];
    else if (lane == 3 * W) mult = -1.0;
    // rows and the reciprocals of their diagonal entries first (independent of the recurrence), then the recurrence itself:
    // one product, one butterfly sum and one multiply per row
    double v[W], nrinv[W];
#pragma unroll
    for (int j = 0; j < W; ++j) v[j] = lane < TW ? T[j * TW + lane] : 0.0;
#pragma unroll
    for (int j = 0; j < W; ++j) nrinv[j] = -fast_rcp(__shfl_sync(FULL, v[j], j));
#pragma unroll
    for (int j = W - 1; j >= 0; --j) {
        double p = v[j] * mult;                               // lanes <= j of the z_M part still hold 0
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) p += __shfl_xor_sync(FULL, p, o);
        if (lane == j) mult = p * nrinv[j];
    }
    if (lane < W) Z[iM * W + lane] = mult;
}
template <int W>
__global__ void __launch_bounds__(128) abc_expand_kernel(const double *TOP, double *Z, int nmerge, int span, int P) {
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= nmerge) return;
    abc_expand_one<W>(TOP, Z, i, span, P, threadIdx.x & 31);
}

// The top of the tree in ONE launch: every level with at most 32 blocks (<= 16 merges; 512 threads keep the merge in registers),
// the root solve and the matching
// expansions, one warp per merge and a CTA barrier between levels (each of those levels would otherwise be a launch of a few warps).
struct AbcTop { int nlev; int cnt[8]; int64_t eoff[9]; int64_t toff[8]; int span0; };
template <int W>
__global__ void __launch_bounds__(512) abc_tree_top_kernel(double *E, double *TOP, double *Z, const AbcTop t, const int P, const int l) {
    __shared__ double Rm[W][W + 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int lv = 0; lv < t.nlev; ++lv) {
        if (w < (t.cnt[lv] + 1) / 2) abc_merge_one<W>(E + t.eoff[lv], E + t.eoff[lv + 1], TOP + t.toff[lv], t.cnt[lv], w, lane);
        __syncthreads();
    }
    if (w == 0) abc_root_one<W>(E + t.eoff[t.nlev], Z, P, l, lane, Rm);
    __syncthreads();
    for (int lv = t.nlev - 1; lv >= 0; --lv) {
        if (w < t.cnt[lv] / 2) abc_expand_one<W>(TOP + t.toff[lv], Z, w, t.span0 << lv, P, lane);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// stage 3: interior back-substitution, one warp per chunk
// ---------------------------------------------------------------------------------------------------------------
// Blocked: LANES = ROWS.  A block of 32 consecutive R rows is staged in shared memory (cp.async, double buffered, so the next
// block streams in while this one is solved); everything that does not depend on the recurrence (separator and tail terms,
// reciprocal of the diagonal) is done by all 32 lanes at once, and the in-block recurrence costs one broadcast + a handful of
// FMAs per row: when x_j is known every lane k < j adds it with coefficient R[k,j] (j - k <= LP, band) or Ufinal[k,:].V[:,j]
// (j - k > LP, fill).
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int LP, int R>
__global__ void __launch_bounds__(128) abc_bs_kernel(const AbcArgs a) {
    using D = AbcDims<LP, R>;
    constexpr int PW = D::PW, NC = D::NC, G0 = D::G0, C0 = D::C0, T0 = D::T0, B0 = D::B0, W = D::W;
    static_assert(NC % 2 == 0, "records are whole 16-byte units");
    constexpr int RQ = R > 0 ? R : 1;
    constexpr int BLK = 32 * NC;                               // doubles per staged block; row stride NC is even, so lane i reading
                                                               // word i*(NC-1)+const is conflict free per half warp
    __shared__ __align__(16) double sm[4][2][BLK];
    __shared__ double smv[4][(32 + LP) * RQ];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int c = blockIdx.x * 4 + wib;
    if (c >= a.P) return;
    const int64_t r0 = chunk_start(a, c);
    const int m = a.m0 + (c < a.jx ? PW : 0), nsteps = m - LP;
    const int64_t r1 = r0 + m;
    const int u = a.u, l = a.l;
    const double *zc = a.Z + (int64_t)c * W, *zn = zc + W;
    const double *v0 = a.V + (r0 + u) * RQ;
    double *xloc = a.x + r0 + u;                               // x of local column 0
    const int64_t xrows = a.n - (r0 + u);                      // local columns that exist in the caller's x
    // separator columns: S_{c+1} = local columns m-LP .. m-1 (global r1-l .. r1+u-1); chunk 0 also owns S_0's real columns
    if (lane < LP) { const int64_t g = r1 - l + lane; if (g >= 0 && g < a.n) a.x[g] = zn[lane]; }
    if (c == 0 && lane < LP) { const int64_t g = r0 - l + lane; if (g >= 0 && g < a.n) a.x[g] = zc[lane]; }
    double zl[W];
#pragma unroll
    for (int t = 0; t < W; ++t) zl[t] = zc[t];
    // xw: lane t < LP holds x of local column kend + t (the LP columns right of the current block); sb = tail sum over columns >= kend + LP
    double xw = lane < LP ? zn[lane] : 0.0;
    double sb[RQ];
#pragma unroll
    for (int q = 0; q < R; ++q) sb[q] = zn[LP + q];
    const double *fr = a.FR + r0 * NC;
    const int nblk = (nsteps + 31) / 32;
    auto prefetch = [&](int bi) {                               // rows past nsteps are never used (their lanes idle); FR is padded
        const double *src = fr + (int64_t)bi * BLK;
        double *dst = sm[wib][bi & 1];
#pragma unroll
        for (int t = 0; t < NC / 2; ++t) cp_async16(dst + 2 * (t * 32 + lane), src + 2 * (t * 32 + lane));
        cp_async_commit();
    };
    prefetch(nblk - 1);
    for (int bi = nblk - 1; bi >= 0; --bi) {
        const int k0 = bi * 32, kend = min(k0 + 32, nsteps), k = k0 + lane;
        // V rows of local columns k0 .. k0+31+LP for the broadcasts below
        __syncwarp();
#pragma unroll
        for (int q = 0; q < R; ++q) {
            smv[wib][lane * R + q] = v0[(int64_t)(k0 + lane) * R + q];
            if (lane < LP) smv[wib][(32 + lane) * R + q] = v0[(int64_t)(k0 + 32 + lane) * R + q];
        }
        if (bi > 0) { prefetch(bi - 1); cp_async_wait<1>(); } else cp_async_wait<0>();
        __syncwarp();
        const double *my = sm[wib][bi & 1] + lane * NC;
        const double *vs = smv[wib];
        double acc = my[B0];
#pragma unroll
        for (int t = 0; t < LP; ++t) acc = fma(-my[C0 + t], zl[t], acc);
#pragma unroll
        for (int q = 0; q < R; ++q) acc = fma(-my[T0 + q], zl[LP + q], acc);
        double g[RQ];
#pragma unroll
        for (int q = 0; q < R; ++q) g[q] = my[G0 + q];
        const double inv = fast_rcp(my[0]);
        // sb is the UNIFORM running tail sum: after column j has been added it covers every column >= j.  Lane k takes its fill term
        // Ufinal[k,:].sb at the moment j = k+LP+1 (everything right of its band); band terms come from its record in shared memory.
        double sbn[RQ];                                          // tail sum over columns >= k0+LP for the next block
#pragma unroll
        for (int q = 0; q < R; ++q) sbn[q] = sb[q];
        auto add_column = [&](const int j, const double xj, const bool tail) {
            const int t = j - k;
            if (tail) {
                double fill = acc;
#pragma unroll
                for (int q = 0; q < R; ++q) { sb[q] = fma(vs[(j - k0) * R + q], xj, sb[q]); fill = fma(-g[q], sb[q], fill); }
                if (t == LP + 1) acc = fill;
            }
            const double band = my[min(max(t, 0), LP)];         // clamped: lanes outside their band read a dummy
            const double coef = (t >= 1 && t <= LP) ? band : 0.0;
            acc = fma(-coef, xj, acc);
        };
        // lanes whose fill starts right of everything seen in this block (k + LP + 1 >= kend + LP) take the incoming tail sum now
        if (k + 1 >= kend) {
#pragma unroll
            for (int q = 0; q < R; ++q) acc = fma(-g[q], sb[q], acc);
        }
        // (a) the LP known columns right of the block, kend+LP-1 .. kend
#pragma unroll
        for (int tt = LP - 1; tt >= 0; --tt) {
            add_column(kend + tt, __shfl_sync(FULL, xw, tt), true);
            if (kend + tt == k0 + LP) {                          // short block (fewer than LP+1 rows): the next block's tail starts here
#pragma unroll
                for (int q = 0; q < R; ++q) sbn[q] = sb[q];
            }
        }
        // (b) the block's own rows, last first: lane j-k0 owns x_j = acc * inv once every later column has been added
        double cand = acc * inv;
        int j = kend - 1;
        for (; j >= k0 + LP; --j) {                              // columns that still feed the tail sum of this or the next block
            add_column(j, __shfl_sync(FULL, cand, j - k0), true);
            cand = acc * inv;
        }
        if (kend - 1 >= k0 + LP) {
#pragma unroll
            for (int q = 0; q < R; ++q) sbn[q] = sb[q];
        }
        for (; j >= k0; --j) {                                   // the first LP rows: band terms only
            add_column(j, __shfl_sync(FULL, cand, j - k0), false);
            cand = acc * inv;
        }
#pragma unroll
        for (int q = 0; q < R; ++q) sb[q] = sbn[q];
        // lanes k <= j never change after their own turn, so cand is x_k
        if (k < kend && k < xrows) xloc[k] = cand;
        // next block's right neighbours = local columns k0 .. k0+LP-1: this block's first rows, then (short block) the old window
        const int nrow = kend - k0;
        const double nx = __shfl_sync(FULL, cand, lane < LP ? lane : 0);
        const double ox = __shfl_sync(FULL, xw, (lane >= nrow && lane < LP) ? lane - nrow : 0);
        xw = lane < LP ? (lane < nrow ? nx : ox) : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// stage 3, many chunks: ONE LANE PER CHUNK over the tiled records.  Row k of a chunk is
//     x_k = ( h_k - C_k . z_sep - T_k . z_tail - sum_{t=1..LP} R[k,k+t] x_{k+t} - Ufinal[k,:] . S_{k+LP+1} ) / R[k,k],
//     S_j = S_{j+1} + V[:,j] x_j  (tail sum over the chunk's columns >= j, started from the separator solution),
// a scalar recurrence whose only serial part is one FMA and one multiply per row; the x window is a ring of LP+1 registers
// indexed at compile time (the sweep is unrolled by LP+1, which divides every chunk's row count).  All 32 lanes do useful work
// on every instruction (the lanes = rows kernel above spends ~35 warp instructions per ROW, this one ~3), so the stage is bound
// by streaming the records: NC x 32 contiguous bytes per row and mini-tile of 4 chunks, pulled through a per-warp bulk-async ring
// (a warp = 32 chunks = 8 mini-tiles).  The V rows
// of the tail-sum update are the one gather (R doubles per lane and row, consecutive rows adjacent in memory): they do not depend
// on the recurrence, so a block's rows are loaded one block (LP+1 rows) ahead.
// ---------------------------------------------------------------------------------------------------------------
// The ring that feeds it: a stage holds ONE row of the warp's eight mini-tiles (8 x NC x 4 doubles); eight lanes issue one
// bulk-async copy each (NC*32 bytes, contiguous in the mini-tile's stream) onto the stage's mbarrier.  Mini-tile blocks are
// padded by 4 doubles in shared memory so that a half warp's 8-byte reads fall into distinct banks.
template <int NC, bool W32 = false>
struct MiniTileRing {
    // W32: records interleaved over the 32 chunks of the tile, FR[tile][k][field][32] (what the lane-per-chunk stage 1 writes: a field of
    // one step is one coalesced 256-byte store for it, and one bulk copy per row here) instead of over mini-tiles of 4 chunks
    static constexpr int MT_STRIDE = NC * 4 + 4, STAGE_DOUBLES = W32 ? NC * 32 : 8 * MT_STRIDE;
    static constexpr uint32_t ROW_BYTES = NC * 32;
    static __host__ __device__ size_t smem_bytes(int nstage) { return (size_t)nstage * STAGE_DOUBLES * 8 + (size_t)nstage * 8; }
    const double *src;            // lanes 0..7: row 0 of mini-tile `lane` of this tile
    double *ring; uint32_t bar0; int nstage, lane; int64_t nsteps, last;
    int cur = 0; uint32_t parity = 0;
    // step g reads row last - g
    __device__ __forceinline__ void issue(int64_t g, int s) {
        const uint32_t bar = bar0 + 8u * s;
        if (W32) {
            if (lane == 0) {
                mbar_expect_tx(bar, 8 * ROW_BYTES);
                bulk_g2s(smem_u32(ring + (size_t)s * STAGE_DOUBLES), src + (last - g) * NC * 32, 8 * ROW_BYTES, bar);
            }
            return;
        }
        if (lane == 0) mbar_expect_tx(bar, 8 * ROW_BYTES);
        __syncwarp();
        if (lane < 8) bulk_g2s(smem_u32(ring + (size_t)s * STAGE_DOUBLES + lane * MT_STRIDE), src + (last - g) * NC * 4, ROW_BYTES, bar);
    }
    __device__ __forceinline__ void init(const double *fr, int64_t tile, int64_t npf, void *smem, int nst, int64_t total_steps) {
        lane = threadIdx.x & 31; nstage = nst; nsteps = total_steps; last = total_steps - 1;
        src = W32 ? fr + tile * npf * NC * 32 : fr + (tile * 8 + (lane & 7)) * npf * NC * 4;
        ring = reinterpret_cast<double *>(smem);
        bar0 = smem_u32(reinterpret_cast<unsigned char *>(smem) + (size_t)nst * STAGE_DOUBLES * 8);
        if (lane == 0) {
            for (int s = 0; s < nst; ++s) mbar_init(bar0 + 8u * s, 1);
            fence_barrier_init();
        }
        __syncwarp();
        for (int g = 0; g < nst && g < total_steps; ++g) issue(g, g);
    }
    // steps must be acquired in order g = 0, 1, 2, ...; rec[f] = field f of this lane's chunk
    __device__ __forceinline__ void acquire(int64_t g, double *rec) {
        const int s = cur;
        mbar_wait(bar0 + 8u * s, parity);
        const double *p = W32 ? ring + (size_t)s * STAGE_DOUBLES + lane : ring + (size_t)s * STAGE_DOUBLES + (lane >> 2) * MT_STRIDE + (lane & 3);
#pragma unroll
        for (int f = 0; f < NC; ++f) rec[f] = p[f * (W32 ? 32 : 4)];
        __syncwarp();                 // every lane has its copy in registers before the stage is re-armed
        fence_proxy_async();          // ... and those reads are ordered before the re-arming bulk copies (see stream_loader.cuh)
        if (g + nstage < nsteps) issue(g + nstage, s);
        if (++cur == nstage) { cur = 0; parity ^= 1u; }
    }
};

template <int LP, int R, bool W32>
__global__ void __launch_bounds__(32) abc_bs_tiled_kernel(const AbcArgs a, const int nstage) {
    using D = AbcDims<LP, R>;
    constexpr int PW = D::PW, NC = D::NC, G0 = D::G0, C0 = D::C0, T0 = D::T0, B0 = D::B0, W = D::W;
    constexpr int RQ = R > 0 ? R : 1;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x;
    const int64_t tile = blockIdx.x;
    int c = (int)(tile * 32 + lane);
    const bool live = c < a.P;
    if (!live) c = a.P - 1;                                    // spare lanes of the last tile shadow chunk P-1 and store nothing
    const int64_t r0 = chunk_start(a, c);
    const int m = a.m0 + (c < a.jx ? PW : 0), nsteps = m - LP;  // the same for every chunk of a tile (abc_layout: jx % 32 == 0 or jx == P)
    const int64_t r1 = r0 + m;
    const int u = a.u, l = a.l;
    const double *zc = a.Z + (int64_t)c * W, *zn = zc + W;
    double *xloc = a.x + r0 + u;                               // x of local column 0
    const int64_t xrows = live ? a.n - (r0 + u) : 0;           // local columns that exist in the caller's x
    // separator columns: S_{c+1} = local columns m-LP .. m-1 (global r1-l .. r1+u-1); chunk 0 also owns S_0's real columns
    if (live) {
#pragma unroll
        for (int t = 0; t < LP; ++t) {
            const int64_t g = r1 - l + t;
            if (g >= 0 && g < a.n) a.x[g] = zn[t];
            const int64_t g0 = r0 - l + t;
            if (c == 0 && g0 >= 0 && g0 < a.n) a.x[g0] = zc[t];
        }
    }
    double zl[W > 0 ? W : 1];
#pragma unroll
    for (int t = 0; t < W; ++t) zl[t] = zc[t];
    // xr[j mod PW] = x of local column j for the PW columns right of the current row; the slot of column k+PW is the one x_k takes
    double xr[PW];
#pragma unroll
    for (int t = 0; t < LP; ++t) xr[t] = zn[t];                // columns nsteps .. nsteps+LP-1 (nsteps is a multiple of PW)
    xr[LP] = 0.0;                                              // column m belongs to the next chunk: it is part of the incoming tail
    double sb[RQ];
#pragma unroll
    for (int q = 0; q < RQ; ++q) sb[q] = R > 0 ? zn[LP + q] : 0.0;
    MiniTileRing<NC, W32> ld;
    ld.init(a.FR, tile, a.npf, smem, nstage, nsteps);
    // vb[ph][:] = V row of local column kb+PW+ph, the column that leaves the band of row kb+ph (rows behind the system are zero)
    const double *vloc = a.V + (r0 + u) * RQ;
    double vb[PW][RQ], vnext[PW][RQ];
    auto load_v = [&](double (&dst)[PW][RQ], const int kb) {
#pragma unroll
        for (int ph = 0; ph < PW; ++ph) {
            const double *src = vloc + (int64_t)(kb + PW + ph) * R;
            if constexpr (R == 2) { const double2 t = __ldg(reinterpret_cast<const double2 *>(src)); dst[ph][0] = t.x; dst[ph][1] = t.y; }
            else {
#pragma unroll
                for (int q = 0; q < R; ++q) dst[ph][q] = __ldg(src + q);
            }
        }
    };
    if (R > 0) load_v(vb, nsteps - PW);
    int64_t g = 0;
    for (int kb = nsteps - PW; kb >= 0; kb -= PW) {
        if (R > 0 && kb >= PW) load_v(vnext, kb - PW);
#pragma unroll
        for (int ph = PW - 1; ph >= 0; --ph) {
            const int k = kb + ph;
            double rec[NC];
            ld.acquire(g++, rec);
            const double xold = xr[ph];                        // x of column k+PW leaves the band: the tail sum now covers columns >= k+LP+1
#pragma unroll
            for (int q = 0; q < R; ++q) sb[q] = fma(vb[ph][q], xold, sb[q]);
            const double inv = fast_rcp(rec[0]);
            double acc = rec[B0];
#pragma unroll
            for (int t = 0; t < LP; ++t) acc = fma(-rec[C0 + t], zl[t], acc);
#pragma unroll
            for (int q = 0; q < R; ++q) acc = fma(-rec[T0 + q], zl[LP + q], acc);
#pragma unroll
            for (int q = 0; q < R; ++q) acc = fma(-rec[G0 + q], sb[q], acc);
#pragma unroll
            for (int t = LP; t >= 1; --t) acc = fma(-rec[t], xr[(ph + t) % PW], acc);     // x_{k+1} last: the only term on the serial chain
            const double x = acc * inv;
            xr[ph] = x;
            if (k < xrows) xloc[k] = x;
        }
        if (R > 0) {
#pragma unroll
            for (int ph = 0; ph < PW; ++ph)
#pragma unroll
                for (int q = 0; q < R; ++q) vb[ph][q] = vnext[ph][q];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// boundary kernels: pack (reference layout -> row records), generator, residual
// ---------------------------------------------------------------------------------------------------------------
// rows n .. n+ABC_PAD-1 are identity rows (the chunked sweep pads the system to a multiple of its unroll period)
__device__ __forceinline__ bool abc_pad_row(double *REC, double *Vd, int64_t i, int64_t n, int l, int u, int r) {
    if (i < n) return false;
    const int lp = l + u, nf = lp + 1 + r;
    for (int d = 0; d < nf; ++d) REC[i * nf + d] = (d == l) ? 1.0 : 0.0;
    for (int q = 0; q < r; ++q) Vd[i * r + q] = 0.0;
    return true;
}
__global__ void abc_pack_kernel(double *REC, double *Vd, const double *bands, const double *U, const double *V, int64_t n, int l, int u, int r) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n + ABC_PAD) return;
    if (abc_pad_row(REC, Vd, i, n, l, u, r)) return;
    const int lp = l + u, nf = lp + 1 + r;
    for (int d = 0; d <= lp; ++d) {
        const int64_t j = i - l + d;
        REC[i * nf + d] = (j >= 0 && j < n) ? bands[(int64_t)(u + i - j) + j * (lp + 1)] : 0.0;
    }
    for (int q = 0; q < r; ++q) { REC[i * nf + lp + 1 + q] = U[i + (int64_t)q * n]; Vd[i * r + q] = V[(int64_t)q + i * r]; }
}
__global__ void abc_generate_kernel(double *REC, double *Vd, int64_t n, int l, int u, int r, uint64_t seed, int dist) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n + ABC_PAD) return;
    if (abc_pad_row(REC, Vd, i, n, l, u, r)) return;
    const uint64_t key = ssm_syskey(seed, 0);
    const int lp = l + u, nf = lp + 1 + r;
    for (int d = 0; d <= lp; ++d) {
        const int64_t j = i - l + d;
        REC[i * nf + d] = (j >= 0 && j < n) ? ssm_ab_band_elem(key, dist, (int)n, l, u, r, lp - d, (int)j) : 0.0;
    }
    for (int q = 0; q < r; ++q) {
        REC[i * nf + lp + 1 + q] = ssm_ab_U_elem(key, dist, (int)n, r, (int)i, q);
        Vd[i * r + q] = ssm_ab_V_elem(key, dist, (int)n, r, q, (int)i);
    }
}
__global__ void abc_vec_generate_kernel(double *b, int64_t n, uint64_t seed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = ssm_ab_b_elem(ssm_syskey(seed, 0), (int)i);
}

// residual: seg sums of V_j x_j, exclusive suffix over segments, then per-row suffix inside each segment
constexpr int RSEG = 2048;     // rows per CTA (256 threads x 8 rows)
__global__ void __launch_bounds__(256) abc_res_partial_kernel(const double *Vd, const double *x, double *part, int64_t n, int r) {
    __shared__ double sh[8][SSM_MAX_RANK];
    const int64_t g0 = (int64_t)blockIdx.x * RSEG;
    double acc[SSM_MAX_RANK] = {0, 0, 0, 0};
    for (int t = threadIdx.x; t < RSEG; t += 256) {
        const int64_t j = g0 + t;
        if (j < n) { const double xj = x[j]; for (int q = 0; q < r; ++q) acc[q] = fma(Vd[j * r + q], xj, acc[q]); }
    }
    for (int q = 0; q < r; ++q) {
        double v = acc[q];
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < r) { double v = 0.0; for (int w = 0; w < 8; ++w) v += sh[w][threadIdx.x]; part[(int64_t)blockIdx.x * SSM_MAX_RANK + threadIdx.x] = v; }
}
__global__ void __launch_bounds__(256) abc_res_scan_kernel(double *part, int nseg, int r, double *acc2) {
    // exclusive suffix scan over the segment sums: thread t owns a contiguous range, block-level suffix in shared memory
    __shared__ double sh[256];
    const int tid = threadIdx.x, per = (nseg + 255) / 256;
    const int g0 = min(nseg, tid * per), g1 = min(nseg, g0 + per);
    for (int q = 0; q < r; ++q) {
        double s = 0.0;
        for (int g = g0; g < g1; ++g) s += part[(int64_t)g * SSM_MAX_RANK + q];
        sh[tid] = s;
        __syncthreads();
        for (int off = 1; off < 256; off <<= 1) {
            const double add = (tid + off < 256) ? sh[tid + off] : 0.0;
            __syncthreads();
            sh[tid] += add;
            __syncthreads();
        }
        double suf = (tid + 1 < 256) ? sh[tid + 1] : 0.0;
        __syncthreads();
        for (int g = g1 - 1; g >= g0; --g) { const double v = part[(int64_t)g * SSM_MAX_RANK + q]; part[(int64_t)g * SSM_MAX_RANK + q] = suf; suf += v; }
    }
    if (tid == 0) { acc2[0] = 0.0; acc2[1] = 0.0; }
}
__global__ void __launch_bounds__(256) abc_res_kernel(const double *REC, const double *Vd, const double *x, const double *b, const double *part,
                                                      double *acc2, int64_t n, int l, int u, int r) {
    __shared__ double sh[256][SSM_MAX_RANK];
    __shared__ double red[2][8];
    const int lp = l + u, nf = lp + 1 + r;
    const int64_t g0 = (int64_t)blockIdx.x * RSEG, g1 = min(n, g0 + RSEG);
    const int tid = threadIdx.x;
    // thread t owns columns g0 + 8t .. g0 + 8t + 7: its sum, then block suffix scan
    double mine[SSM_MAX_RANK] = {0, 0, 0, 0};
    for (int e = 0; e < 8; ++e) { const int64_t j = g0 + tid * 8 + e; if (j < g1) { const double xj = x[j]; for (int q = 0; q < r; ++q) mine[q] = fma(Vd[j * r + q], xj, mine[q]); } }
    for (int q = 0; q < r; ++q) sh[tid][q] = mine[q];
    __syncthreads();
    // suffix (exclusive) over threads, simple log-step scan
    for (int off = 1; off < 256; off <<= 1) {
        double add[SSM_MAX_RANK];
        for (int q = 0; q < r; ++q) add[q] = (tid + off < 256) ? sh[tid + off][q] : 0.0;
        __syncthreads();
        for (int q = 0; q < r; ++q) sh[tid][q] += add[q];
        __syncthreads();
    }
    // sh[tid] = inclusive suffix from thread tid; Suf(k) for k = first column of thread tid+1 is sh[tid+1] + segment suffix
    double suf[SSM_MAX_RANK];
    for (int q = 0; q < r; ++q) suf[q] = ((tid + 1 < 256) ? sh[tid + 1][q] : 0.0) + part[(int64_t)blockIdx.x * SSM_MAX_RANK + q];
    // walk own rows descending; S(i) = Suf(i+u+1).  Start with Suf(k0) where k0 = g0 + 8(tid+1), then add columns down to i+u+1
    double num = 0.0, den = 0.0;
    int64_t kcur = g0 + (int64_t)(tid + 1) * 8;           // suf = Suf(kcur) restricted ... columns >= kcur (within or beyond the segment)
    if (kcur > g1) kcur = g1;
    for (int e = 7; e >= 0; --e) {
        const int64_t i = g0 + tid * 8 + e;
        if (i >= g1) continue;
        const int64_t need = i + u + 1;                   // want Suf(need)
        double s[SSM_MAX_RANK];
        for (int q = 0; q < r; ++q) s[q] = suf[q];
        if (need >= kcur) {                               // remove columns kcur .. need-1
            for (int64_t j = kcur; j < need && j < n; ++j) { const double xj = x[j]; for (int q = 0; q < r; ++q) s[q] = fma(-Vd[j * r + q], xj, s[q]); }
        } else {                                          // add columns need .. kcur-1
            for (int64_t j = need; j < kcur; ++j) { const double xj = x[j]; for (int q = 0; q < r; ++q) s[q] = fma(Vd[j * r + q], xj, s[q]); }
        }
        double ax = 0.0;
        for (int d = 0; d <= lp; ++d) { const int64_t j = i - l + d; if (j >= 0 && j < n) ax = fma(REC[i * nf + d], x[j], ax); }
        for (int q = 0; q < r; ++q) ax = fma(REC[i * nf + lp + 1 + q], s[q], ax);
        const double bi = b[i], dl = bi - ax;
        num = fma(dl, dl, num); den = fma(bi, bi, den);
    }
    for (int o = 16; o >= 1; o >>= 1) { num += __shfl_xor_sync(FULL, num, o); den += __shfl_xor_sync(FULL, den, o); }
    if ((tid & 31) == 0) { red[0][tid >> 5] = num; red[1][tid >> 5] = den; }
    __syncthreads();
    if (tid == 0) { double a0 = 0, a1 = 0; for (int w = 0; w < 8; ++w) { a0 += red[0][w]; a1 += red[1][w]; } atomicAdd(&acc2[0], a0); atomicAdd(&acc2[1], a1); }
}

}  // namespace ssm

// =================================================================================================================
// C ABI
// =================================================================================================================
using namespace ssm;

struct ssm_abc {
    ssm_ctx *ctx; ssm::CtxRef ref; int64_t n; int l, u, r, lp, w, nf, nc;
    DevBufPtr REC, V;                 // operands
    DevBufPtr FR, E, TOP, Z, tmp;     // workspace of the last solve (allocated on first use)
    DevBufPtr RI, VI;                 // chunk-interleaved entering rows of the lane-per-chunk stage 1, valid for the operands and the layout below
    int ri_P = 0, ri_m0 = 0, ri_jx = 0;
    int nsolve = 0;                   // solves since the operands were written
    int P = 0, m0 = 0, jx = 0;
    bool tiled = false;               // layout of FR in the last solve
    int64_t npf = 0;
};

#define SSM_CHECK_ARG(cond, msg) do { if (!(cond)) { set_error("%s: %s", __func__, msg); return SSM_E_ARG; } } while (0)

#define SSM_ABC_SHAPES(X) X(1, 0) X(2, 0) X(4, 0) X(1, 1) X(1, 2) X(2, 1) X(2, 2) X(3, 1) X(3, 2) X(4, 2) X(6, 2) X(8, 2)

static bool abc_supported(int lp, int r) {
#define X(LP_, R_) if (lp == LP_ && r == R_) return true;
    SSM_ABC_SHAPES(X)
#undef X
    return false;
}

// Device views of up to four caller arrays: device pointers pass through, host arrays share ONE carve-up of the context's
// grow-only staging buffer (no cudaMalloc / cudaFree per call: for a single system of n <= 10^6 those cost more than the solve).
// `copy_in[k]` = the array is an input (copied host -> device on the context's stream).
struct Staged {
    const double *host[4]; size_t count[4]; bool copy_in[4]; double *dev[4]; bool staged[4]; int n = 0;
    int add(const double *p, size_t cnt, bool in) { host[n] = p; count[n] = cnt; copy_in[n] = in; dev[n] = nullptr; staged[n] = false; return n++; }
};
static int stage_arrays(ssm_ctx *c, Staged &st) {
    size_t total = 0;
    for (int k = 0; k < st.n; ++k) {
        if (!st.host[k]) continue;
        if (is_device_ptr(st.host[k])) { st.dev[k] = const_cast<double *>(st.host[k]); continue; }
        st.staged[k] = true;
        total += (st.count[k] + 1) & ~(size_t)1;               // keep every carve-out 16-byte aligned
    }
    if (total == 0) return 0;
    double *base;
    SSM_TRY(staging_dev(c, total, &base));
    for (int k = 0; k < st.n; ++k) {
        if (!st.staged[k]) continue;
        st.dev[k] = base;
        base += (st.count[k] + 1) & ~(size_t)1;
        if (st.copy_in[k]) SSM_CUDA(cudaMemcpyAsync(st.dev[k], st.host[k], st.count[k] * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    return 0;
}

// regular chunk layout for P chunks: base length m0 = pw*q - 1, the first jx chunks get pw more rows; P is reduced until q >= 2
// `same` = how many consecutive chunks must share a length: the 4 chunks of a stage-1 warp, or the 32 of a stage-3 tile
static void abc_layout(int64_t n, int lp, int same, int &P, int &m0, int &jx) {
    const int pw = lp + 1;
    if (P < 1) P = 1;
    while (P > 1 && (n / P + 1) / pw < 2) --P;
    int64_t q = (n / P + 1) / pw;
    if (q < 2) q = 2;                                        // tiny n: a single padded chunk
    m0 = (int)(pw * q - 1);
    const int64_t rem = n - (int64_t)P * m0;
    jx = rem > 0 ? (int)((rem + pw - 1) / pw) : 0;
    jx = ((jx + same - 1) / same) * same;
    if (jx > P) jx = P;
}

// shapes with a lane-per-chunk instance of stage 1
constexpr bool abc_has_lpc(int lp, int r) { return r == 2 && (lp == 8 || lp == 6 || lp == 4); }
// shapes with a 4-lanes-per-chunk instance of stage 1 (the wide active blocks), and the group size a context asks for
constexpr bool abc_has_gs4(int lp, int r) { return 2 * lp + 2 * r + 2 >= 18; }
static int abc_gs_of(const ssm_ctx *c, int lp, int r) { return abc_has_gs4(lp, r) && c->abc_gs == 4 ? 4 : 8; }

static int abc_default_chunks(const ssm_abc *A) {
    // Stage 1 holds 12 warps of 4 chunks per SM, and all chunks are equally long, so its time moves in whole WAVES of
    // sm_count * 48 chunks.  Large systems: ~512 rows per chunk, rounded down to whole waves, at most 4 (measured at n = 2^24:
    // 28,416 and 32,768 chunks 2.89 ms, 35,520 2.96, 42,624 3.14).  Below one wave of 512-row chunks the sweep is latency bound and
    // shorter chunks win until the separator tree takes over: 128 rows per chunk, at most one wave (n = 10^6: 1,953 chunks 0.50 ms,
    // 7,104 0.40 ms; n = 65,536: 128 chunks 0.37 ms, 512 0.20 ms).
    // (4 lanes per chunk: 8 warps of 8 chunks per SM, waves of sm_count * 64; n = 2^24 gets 28,416 chunks either way)
    // One lane per chunk ((l+u, r) = (8, 2)): 3 CTAs of 32 chunks per SM, waves of sm_count * 96 = 14,208 chunks — 28,416 at n = 2^24 once
    // more.  It needs tiled records and wins from one full wave of >= 128-row chunks on (measured, n = 2 / 3 / 4 / 6 M: 0.47 / 0.59 / 0.71 /
    // 0.95 ms with 14,208 chunks against 0.50 / 0.65 / 0.79 / 1.10 ms with the lanes = columns kernels at 9,472).
    const int64_t sms = A->ctx->sm_count > 0 ? A->ctx->sm_count : 148;
    const int64_t wave_lpc = sms * (A->lp == 4 ? 5 : 3) * 32;      // CTAs per SM: 168 / 138 registers at l+u = 8 / 6, 96 at l+u = 4
    const bool lpc = abc_has_lpc(A->lp, A->r) && A->ctx->abc_lpc && A->n / 128 >= wave_lpc && wave_lpc >= A->ctx->abc_tiled_min_chunks;
    const int per_sm = lpc ? (int)(wave_lpc / sms) : abc_gs_of(A->ctx, A->lp, A->r) == 4 ? 8 * 8 : 12 * ABC_CPW;
    const int64_t wave = sms * per_sm;
    int64_t P = A->n / 512;
    if (lpc) P = std::min<int64_t>(std::max<int64_t>(A->n / 1024 / wave, 1), 2) * wave;   // ~1,024 rows per chunk, one or two waves (measured:
                                                                                          // n = 2^24 14,208 / 28,416 chunks 2.02 / 2.04 ms,
                                                                                          // 2^25 3.87 / 3.77 (56,832: 3.89), 10^8 10.99 / 10.61 (56,832: 10.73))
    else if (P >= wave) P = std::min<int64_t>(P / wave, 4) * wave;
    else P = std::min<int64_t>(A->n / 128, wave);
    const int64_t pmax = A->n / (2 * A->lp + 2 + 8);
    if (P > pmax) P = pmax;
    if (P < 1) P = 1;
    return (int)P;
}

template <int LP, int R>
static int abc_solve_inst(ssm_abc *A, const double *b, double *x, int P) {
    ssm_ctx *c = A->ctx;
    using D = AbcDims<LP, R>;
    constexpr int W = D::W, EW = 2 * W + 1, TW = 3 * W + 1;
    AbcArgs a{A->REC->p, A->V->p, b, A->FR->p, A->E->p, A->Z->p, x, A->n, A->l, A->u, P, A->m0, A->jx, LP + 1, A->npf,
              reinterpret_cast<int *>(A->Z->p + ((int64_t)P + 2) * W), nullptr, nullptr, 0};
    const unsigned gchunks = (unsigned)((P + 3) / 4);
    // lanes per chunk in stage 1: 8 (4 chunks per warp) or, for the wide active blocks, 4 (8 chunks per warp; option "abc_gs")
    auto stage1 = [&](auto gs_tag) -> int {
        constexpr int GS = decltype(gs_tag)::value, CPW = 32 / GS;
        const unsigned gqr = (unsigned)((P + 4 * CPW - 1) / (4 * CPW));
        constexpr int smem = AbcQrSmem<LP, R, GS>::TOTAL;
        auto launch = [&](auto kern) -> int {
            if (smem > 48 * 1024) SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            kern<<<gqr, 128, smem, c->stream>>>(a);
            return 0;
        };
        if (A->tiled) SSM_TRY(launch(abc_qr_kernel<LP, R, GS, true>));
        else SSM_TRY(launch(abc_qr_kernel<LP, R, GS, false>));
        return 0;
    };
    bool lpc_done = false;
    if constexpr (abc_has_lpc(LP, R)) {
        // one lane per chunk over tiles of 32 equally long chunks.  It sweeps an interleaved copy of the operands that costs a pass over
        // them to build: operands that are solved against once (the Julia glue's `A \ b`) never pay for it — option abc_lpc = 1 switches
        // over at the SECOND solve of the same operands, 2 at the first, 0 never
        if (A->tiled && (c->abc_lpc >= 2 || (c->abc_lpc == 1 && (A->RI || A->nsolve >= 1)))) {
            using LS = AbcLpcSmem<LP, R>;
            const int64_t ntile = (P + 31) / 32;
            const int nrow = A->m0 + (LP + 1) - LP;            // steps of a long chunk
            if (!A->RI || A->ri_P != P || A->ri_m0 != A->m0 || A->ri_jx != A->jx) {      // (operand setters drop RI: see ssm_abc_set / _generate)
                A->RI.reset(); A->VI.reset();
                SSM_TRY(dev_alloc(c, (size_t)ntile * nrow * D::NF * 32, false, A->RI));
                SSM_TRY(dev_alloc(c, (size_t)ntile * nrow * R * 32, false, A->VI));
                abc_interleave_kernel<D::NF, R><<<dim3((unsigned)((nrow + 7) / 8), (unsigned)ntile), 256, 0, c->stream>>>(a, A->RI->p, A->VI->p, LP, nrow);
                c->launches++;
                A->ri_P = P; A->ri_m0 = A->m0; A->ri_jx = A->jx;
            }
            a.RI = A->RI->p; a.VI = A->VI->p; a.nrow_i = nrow;
            auto kern = abc_qr_lpc_kernel<LP, R, true>;
            SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LS::TOTAL));
            SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            kern<<<(unsigned)ntile, 128, LS::TOTAL, c->stream>>>(a);
            lpc_done = true;
        }
    }
    if (lpc_done) {
    } else if constexpr (abc_has_gs4(LP, R)) {
        if (abc_gs_of(c, LP, R) == 4) SSM_TRY(stage1(std::integral_constant<int, 4>{}));
        else SSM_TRY(stage1(std::integral_constant<int, 8>{}));
    } else {
        SSM_TRY(stage1(std::integral_constant<int, 8>{}));
    }
    c->launches++;
    // levels: level t has cnt[t] blocks stored at E + eoff[t]; levels with more than 32 blocks are one launch each, the rest of the
    // tree (merges, root, expansions) is a single launch
    std::vector<int> cnt; std::vector<int64_t> eoff, toff;
    int cur = P; int64_t eo = 0, to = 0;
    cnt.push_back(cur); eoff.push_back(0);
    while (cur > 1) {
        const int nout = (cur + 1) / 2;
        const int64_t nexto = eo + (int64_t)cur * W * EW;
        toff.push_back(to);
        to += (int64_t)(cur / 2) * W * TW;
        eo = nexto; cur = nout;
        cnt.push_back(cur); eoff.push_back(eo);
    }
    const int nlevels = (int)toff.size();
    int t0 = 0;
    while (t0 < nlevels && cnt[t0] > 32) ++t0;                 // first level handled by the fused top kernel
    for (int t = 0; t < t0; ++t) {
        abc_merge_kernel<W><<<(unsigned)(((cnt[t] + 1) / 2 + 3) / 4), 128, 0, c->stream>>>(A->E->p + eoff[t], A->E->p + eoff[t + 1], A->TOP->p + toff[t], cnt[t]);
        c->launches++;
    }
    {
        AbcTop top{};
        top.nlev = nlevels - t0; top.span0 = 1 << t0;
        for (int k = 0; k < top.nlev; ++k) { top.cnt[k] = cnt[t0 + k]; top.eoff[k] = eoff[t0 + k]; top.toff[k] = toff[t0 + k]; }
        top.eoff[top.nlev] = eoff[nlevels];
        abc_tree_top_kernel<W><<<1, 512, 0, c->stream>>>(A->E->p, A->TOP->p, A->Z->p, top, P, A->l);
        c->launches++;
    }
    for (int t = t0 - 1; t >= 0; --t) {
        const int nmerge = cnt[t] / 2;
        if (nmerge == 0) continue;
        abc_expand_kernel<W><<<(unsigned)((nmerge + 3) / 4), 128, 0, c->stream>>>(A->TOP->p + toff[t], A->Z->p, nmerge, 1 << t, P);
        c->launches++;
    }
    if (A->tiled) {
        // ring depth: the shared memory an SM has, shared by the single-warp CTAs the grid puts on it (at most 16 stages)
        auto stage3 = [&](auto w32_tag) -> int {
            constexpr bool W32 = decltype(w32_tag)::value;       // records interleaved over 32 chunks (written by the lane-per-chunk stage 1) or over 4
            using LD = MiniTileRing<D::NC, W32>;
            const int64_t ntiles = (P + 31) / 32;
            int64_t per_sm = (ntiles + c->sm_count - 1) / (c->sm_count > 0 ? c->sm_count : 148);
            per_sm = std::min<int64_t>(std::max<int64_t>(per_sm, 1), 16);
            int ns = (int)((size_t)(208 * 1024) / (size_t)per_sm / (LD::STAGE_DOUBLES * 8 + 8));
            ns = std::min(std::max(ns, 2), 16);
            if (c->pipeline_stages > 0) ns = c->pipeline_stages;
            const size_t smem = LD::smem_bytes(ns);
            auto kern = abc_bs_tiled_kernel<LP, R, W32>;
            if (smem > 48 * 1024) SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            kern<<<(unsigned)ntiles, 32, smem, c->stream>>>(a, ns);
            return 0;
        };
        if (lpc_done) SSM_TRY(stage3(std::true_type{}));
        else SSM_TRY(stage3(std::false_type{}));
    } else {
        abc_bs_kernel<LP, R><<<gchunks, 128, 0, c->stream>>>(a);
    }
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}

// A context keeps the last two destroyed objects of moderate size (operands + workspace of one system, <= ~300 MB each): a caller
// that solves one system after another of the same shape (`A \ b` from the Julia glue creates and destroys an object per call)
// then pays no cudaMalloc / cudaFree / memset per solve.  Every entry point overwrites what it reads, so a recycled object needs no
// clearing: ssm_abc_set / ssm_abc_generate write all n + ABC_PAD operand rows, a solve writes every workspace entry it later reads.
constexpr int64_t ABC_CACHE_MAX_N = 1 << 20;
void ssm::abc_cache_free(ssm_ctx *c) {
    for (auto &slot : c->abc_cache) { delete static_cast<ssm_abc *>(slot); slot = nullptr; }
}

extern "C" {

int ssm_abc_create(ssm_ctx *c, int64_t n, int l, int u, int r, ssm_abc **out) {
    SSM_CHECK_ARG(c && out, "null argument");
    SSM_CHECK_ARG(n >= 1 && l >= 0 && u >= 0 && r >= 0, "bad sizes");
    if (n > (int64_t)0x7fffffff - 1024) { set_error("ssm_abc_create: n too large"); return SSM_E_UNSUPPORTED; }
    if (!abc_supported(l + u, r)) { set_error("ssm_abc: (l+u, r) = (%d, %d) has no compiled chunked kernel", l + u, r); return SSM_E_UNSUPPORTED; }
    for (auto &slot : c->abc_cache) {
        auto *C = static_cast<ssm_abc *>(slot);
        if (C && C->n == n && C->l == l && C->u == u && C->r == r) { slot = nullptr; *out = C; return 0; }
    }
    auto *A = new ssm_abc();
    A->ctx = c; A->ref.bind(c); A->n = n; A->l = l; A->u = u; A->r = r; A->lp = l + u; A->w = A->lp + r; A->nf = A->lp + 1 + r; A->nc = 2 * A->lp + 2 * r + 2;
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(c->device);
    int rc = dev_alloc(c, (size_t)(n + ABC_PAD) * A->nf, true, A->REC);
    if (!rc) rc = dev_alloc(c, (size_t)(n + ABC_PAD) * (r > 0 ? r : 1), true, A->V);
    cudaSetDevice(prev);
    if (rc) { delete A; return rc; }
    *out = A;
    return 0;
}
int ssm_abc_destroy(ssm_abc *A) {
    if (!A) return 0;
    ssm_ctx *c = A->ctx;
    obj_sync(c);
    if (A->n <= ABC_CACHE_MAX_N && !c->retired.load()) {
        A->tmp.reset();
        if (c->abc_cache[0]) { delete static_cast<ssm_abc *>(c->abc_cache[1]); c->abc_cache[1] = c->abc_cache[0]; }
        c->abc_cache[0] = A;
        return 0;
    }
    delete A;
    return 0;
}

int ssm_abc_generate(ssm_abc *A, uint64_t seed, int dist) {
    SSM_CHECK_ARG(A, "null argument");
    ssm_ctx *c = A->ctx;
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(c->device);
    A->RI.reset(); A->VI.reset(); A->nsolve = 0;              // the interleaved copy belongs to the old operands
    abc_generate_kernel<<<(unsigned)((A->n + ABC_PAD + 255) / 256), 256, 0, c->stream>>>(A->REC->p, A->V->p, A->n, A->l, A->u, A->r, seed, dist);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    cudaSetDevice(prev);
    if (e != cudaSuccess) return cuda_fail(e, "abc_generate_kernel", __FILE__, __LINE__);
    return 0;
}

int ssm_abc_set(ssm_abc *A, const double *bands, const double *U, const double *V) {
    SSM_CHECK_ARG(A && bands && (A->r == 0 || (U && V)), "null argument");
    ssm_ctx *c = A->ctx;
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(c->device);
    A->RI.reset(); A->VI.reset(); A->nsolve = 0;              // the interleaved copy belongs to the old operands
    Staged st;
    const int kb = st.add(bands, (size_t)(A->lp + 1) * A->n, true), ku = st.add(A->r ? U : nullptr, (size_t)A->n * A->r, true),
              kv = st.add(A->r ? V : nullptr, (size_t)A->n * A->r, true);
    int rc = stage_arrays(c, st);
    if (!rc) {
        abc_pack_kernel<<<(unsigned)((A->n + ABC_PAD + 255) / 256), 256, 0, c->stream>>>(A->REC->p, A->V->p, st.dev[kb], st.dev[ku], st.dev[kv], A->n, A->l, A->u, A->r);
        c->launches++;
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "abc_pack_kernel", __FILE__, __LINE__);
    }
    cudaSetDevice(prev);
    return rc;
}

int ssm_abc_vec_generate(ssm_abc *A, double *b_dev, uint64_t seed) {
    SSM_CHECK_ARG(A && b_dev && is_device_ptr(b_dev), "b must be a device pointer");
    ssm_ctx *c = A->ctx;
    abc_vec_generate_kernel<<<(unsigned)((A->n + 255) / 256), 256, 0, c->stream>>>(b_dev, A->n, seed);
    c->launches++;
    SSM_CUDA(cudaGetLastError());
    return 0;
}

// A \ b for one large system: x <- A \ b (b, x: n doubles, host or device; x may alias b).  nchunks = 0 picks a default.
int ssm_abc_solve(ssm_abc *A, const double *b, double *x, int nchunks) {
    SSM_CHECK_ARG(A && b && x, "null argument");
    ssm_ctx *c = A->ctx;
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(c->device);
    int P = nchunks > 0 ? nchunks : abc_default_chunks(A);
    int m0 = 0, jx = 0;
    abc_layout(A->n, A->lp, 32 / abc_gs_of(c, A->lp, A->r), P, m0, jx);
    const bool tiled = P >= c->abc_tiled_min_chunks;          // P only shrinks for tiny n, where neither rounding changes it
    if (tiled) abc_layout(A->n, A->lp, 32, P, m0, jx);
    const int64_t npf = (int64_t)m0 + 1;                      // rows eliminated by a long chunk: m0 + (lp+1) - lp
    int rc = 0;
    // rows the sweeps may touch behind the padded system: one more block of lp+1 operand rows (stage-1 ring, fetched ahead at the
    // last step) and the V rows of u + lp + 2 further columns, plus the 16-byte slack of the bulk copies
    if ((int64_t)P * m0 + (int64_t)jx * (A->lp + 1) > A->n + ABC_PAD - (2 * (A->lp + 1) + A->u + 8)) { set_error("ssm_abc_solve: internal chunk layout exceeds the padding"); rc = SSM_E_ARG; }
    const int W = A->w;
    if (!rc && (!A->FR || A->P != P || A->tiled != tiled || A->npf != npf)) {
        A->FR.reset(); A->E.reset(); A->TOP.reset(); A->Z.reset();
        const size_t frcount = tiled ? (size_t)((P + 31) / 32) * (size_t)npf * (size_t)A->nc * 32 : (size_t)(A->n + ABC_PAD + 40) * A->nc;
        rc = dev_alloc(c, frcount, tiled, A->FR);   // tiled: spare lanes of the last tile are read, so they are defined
        if (!rc) rc = dev_alloc(c, (size_t)(2 * (int64_t)P + 64) * W * (2 * W + 1), true, A->E);
        if (!rc) rc = dev_alloc(c, (size_t)((int64_t)P + 64) * W * (3 * W + 1), true, A->TOP);
        if (!rc) rc = dev_alloc(c, (size_t)((int64_t)P + 2) * W + 256, true, A->Z);   // (+ 512 per-SM arrival counters of the lane-per-chunk stage 1)
        A->P = P;
    }
    A->m0 = m0; A->jx = jx; A->tiled = tiled; A->npf = npf;
    Staged st;
    const int kb = st.add(b, (size_t)A->n, true), kx = st.add(x, (size_t)A->n, false);
    if (!rc) rc = stage_arrays(c, st);
    const double *db = st.dev[kb]; double *dx = st.dev[kx];
    const bool xhost = st.staged[kx];
    if (!rc) {
        rc = SSM_E_UNSUPPORTED;
#define X(LP_, R_) if (A->lp == LP_ && A->r == R_) rc = abc_solve_inst<LP_, R_>(A, db, dx, P);
        SSM_ABC_SHAPES(X)
#undef X
        if (!rc) A->nsolve++;
    }
    if (!rc && xhost) {
        cudaError_t e = cudaMemcpyAsync(x, dx, (size_t)A->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "D2H x", __FILE__, __LINE__);
    }
    if (!rc && (xhost || st.staged[kb])) {
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
    }
    cudaSetDevice(prev);
    return rc;
}

// relres = ||b - A x||_2 / ||b||_2 with A applied through its structure (AlmostBandedMatrix.jl:84-90)
int ssm_abc_residual(ssm_abc *A, const double *x, const double *b, double *relres) {
    SSM_CHECK_ARG(A && x && b && relres, "null argument");
    ssm_ctx *c = A->ctx;
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(c->device);
    Staged st;
    const int kx = st.add(x, (size_t)A->n, true), kb = st.add(b, (size_t)A->n, true);
    int rc = stage_arrays(c, st);
    const double *dx = st.dev[kx], *db = st.dev[kb];
    const int nseg = (int)((A->n + RSEG - 1) / RSEG);
    if (!rc) rc = dev_alloc(c, (size_t)nseg * SSM_MAX_RANK + 2, true, A->tmp);
    if (!rc) {
        double *part = A->tmp->p, *acc2 = part + (size_t)nseg * SSM_MAX_RANK;
        abc_res_partial_kernel<<<nseg, 256, 0, c->stream>>>(A->V->p, dx, part, A->n, A->r);
        abc_res_scan_kernel<<<1, 256, 0, c->stream>>>(part, nseg, A->r, acc2);
        abc_res_kernel<<<nseg, 256, 0, c->stream>>>(A->REC->p, A->V->p, dx, db, part, acc2, A->n, A->l, A->u, A->r);
        c->launches += 3;
        double h[2] = {0, 0};
        cudaError_t e = cudaMemcpyAsync(h, acc2, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "abc residual", __FILE__, __LINE__);
        else *relres = sqrt(h[0]) / sqrt(h[1]);
    }
    cudaSetDevice(prev);
    return rc;
}

// inspection (tests): copy the level-0 reduced blocks / separator solution / R records of the last solve to the host
int ssm_abc_debug_get(ssm_abc *A, int what, double *out, int64_t count) {
    SSM_CHECK_ARG(A && out, "null argument");
    DevBufPtr src = what == 0 ? A->E : what == 1 ? A->Z : what == 2 ? A->FR : A->REC;
    if (!src) { set_error("ssm_abc_debug_get: no solve has run"); return SSM_E_STATE; }
    if ((size_t)count > src->count) count = (int64_t)src->count;
    SSM_CUDA(cudaMemcpyAsync(out, src->p, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, A->ctx->stream));
    SSM_CUDA(cudaStreamSynchronize(A->ctx->stream));
    return 0;
}
int ssm_abc_chunks(const ssm_abc *A) { return A ? A->P : 0; }

}  // extern "C"
