// Persistent recurrent part of one LSTM layer on the 5th-generation tensor cores (hac: H = 384).
// Reference semantics: bonito/nn.py:353-415 (torch.nn.LSTM, gate order i,f,g,o, zero initial state, optional
// time reversal), the span `Model.use_koi` hands to koi.lstm (bonito/crf/model.py:240-246).
//
// Decomposition: a cluster of CS = 8 CTAs owns one batch tile of NB = 32 chunks for all T steps; CTA `rank` owns
// hidden units [48*rank, 48*rank+48) = 192 gate rows of W_hh.
//   * W_hh slice lives in TENSOR MEMORY for the whole kernel as the UMMA A operand: two 128-lane blocks
//     (rows 0..127 and rows 64..191; fp16 pairs per 32-bit column => 192 columns each).  The tensor core re-reads
//     all of it every step; from shared memory that costs twice as long (59 vs 30 cycles per 128x32x16 MMA measured).
//   * the tile is processed as TWO interleaved sub-tiles of 16 chunks: while the epilogue warps run the cell update and
//     the exchange of sub-tile A, the tensor core already works on sub-tile B.  The cost of an MMA here is set by the
//     A-operand read, not by N, so the split is free on the tensor side and hides ~1/3 of the per-step chain.
//   * h_{t-1} of a sub-tile [16 chunks x 384] sits in shared memory as the UMMA B operand, K-major WITHOUT swizzle:
//     [48 k-chunks of 8 units][16 chunks][16 B], so the 8 units x 16 chunks one warp produces are 256 contiguous
//     bytes of every peer's tile (double buffered by step parity).
//   * per (step, sub-tile) one elected thread issues 2 x 24 tcgen05.mma (M=128, N=16, K=16) -> gate pre-activations in
//     TMEM; eight epilogue warps PER SUB-TILE (two independent sets, so the two sub-tiles' cell updates overlap in time
//     instead of alternating on the same warps: 2412 instead of 2811 cycles per step) pull them with tcgen05.ld.16x256b -- the mma-accumulator fragment, so with rows
//     ordered [8 units x (i,f,g,o)] one thread holds all four gates of a (unit, chunk) -- add the prefetched input
//     projection, update (c, h) in registers, stage the new h block in shared memory and push it into the h tile
//     of all 8 CTAs of the cluster with one bulk copy per peer (cp.async.bulk shared::cta -> shared::cluster; the copy
//     completes transaction bytes on the destination's mbarrier, so there is no fence, no arrive and no per-lane
//     remote store on the sender), and write it to Y[t].  No __threadfence, no cluster barrier, no L2 round trip.
//
// Packed operands are the same as for the mma.sync kernel (lstm_rec.cu): whh [CS][UPC/8][gate][8][H],
// gx [T][N][CS][UPC/8][8][gate], y [T][N][H].
#include <stdlib.h>

#include "tc_common.cuh"

namespace {

constexpr int NB = 32;             // chunks per cluster
constexpr int NS = 2;              // interleaved sub-tiles
constexpr int SN = NB / NS;        // 16 chunks per sub-tile = N of one MMA
constexpr int H = 384;
constexpr int CS = 8;
constexpr int UPC = H / CS;        // 48 units per CTA
constexpr int ROWS = 4 * UPC;      // 192 gate rows per CTA
constexpr int THREADS = 288;       // 8 epilogue warps + the MMA warp
constexpr int THREADS_SPLIT = 544; // 2 x 8 epilogue warps (one set per sub-tile) + the MMA warp
constexpr uint32_t HT = (H / 8) * SN * 16;         // one h tile: 48 k-chunks x 16 chunks x 16 B = 12288 B
constexpr uint32_t COL_A1 = 0, COL_A2 = 192, COL_D = 384, TMEM_COLS = 512;   // D: [sub][D1|D2] x 16 columns
constexpr uint32_t STAGE_WARP = SN * 16;           // 256 B per (parity, sub, warp)
constexpr uint32_t OFF_H = 0;                      // [sub][parity] h tiles
constexpr uint32_t OFF_STAGE = NS * 2 * HT;        // [parity][sub][warp]
constexpr uint32_t OFF_BARS = OFF_STAGE + 2 * NS * 8 * STAGE_WARP;
constexpr uint32_t SMEM_USED = OFF_BARS + 128 + 1024;
// the kernel owns all 512 TMEM columns: ask for more than half of the SM's shared memory so that two CTAs can
// never be co-resident (a second tcgen05.alloc on the same SM would spin forever)
constexpr uint32_t SMEM_BYTES = SMEM_USED > 120 * 1024 ? SMEM_USED : 120 * 1024;

// timeline of CTA 0 (VARIANT 3), sub-tile 0: per step, SM-clock stamps of
//   [0] h tile complete (MMA thread)   [1] MMAs issued + committed   [2] accumulator ready (epilogue warp 0)
//   [3] TMEM loaded   [4] cell update done   [5] h chunk sent        [6] %globaltimer (ns) at [0]     [7] [5] for warp 7
constexpr int TL_STEPS = 256;
__device__ long long g_timeline[TL_STEPS][8];

// sigma(i), sigma(f), tanh(g), sigma(o) from four ex2 and ONE reciprocal (batch inversion); the exponent arguments are
// clamped so the product of the four denominators stays finite (sigma(-20.8) = 9e-10: the clamp is invisible in fp16).
__device__ __forceinline__ void gate_activations(float ai, float af, float ag, float ao, float& si, float& sf, float& tg,
                                                 float& so) {
    constexpr float L = 1.4426950408889634f, CLAMP = 30.0f;
    const float di = 1.0f + ex2_approx(fminf(-L * ai, CLAMP));
    const float df = 1.0f + ex2_approx(fminf(-L * af, CLAMP));
    const float dg = 1.0f + ex2_approx(fminf(-2.0f * L * ag, CLAMP));
    const float dO = 1.0f + ex2_approx(fminf(-L * ao, CLAMP));
    const float pif = di * df, pgo = dg * dO;
    const float r = rcp_approx(pif * pgo);
    const float rif = r * pgo, rgo = r * pif;
    si = rif * df;
    sf = rif * di;
    tg = fmaf(2.0f, rgo * dO, -1.0f);
    so = rgo * dg;
}

struct RecBars {
    uint32_t hfull;   // [sub][parity] at hfull + 8*(2*sub + parity)
    uint32_t dfull;   // [sub] at dfull + 8*sub
};

// One epilogue warp: row block `blk` (8 hidden units x 4 gates = 32 TMEM lanes at lane quarter `quarter`), chunks
// col0 .. col0 + 8*NJ - 1 of EACH 16-chunk sub-tile, accumulator `which` (0: rows 0..127, 1: rows 64..191).
// VARIANT is a timing-experiment knob (B200_LSTM_DEBUG): 0 = product; 1 = all eight copies of the h block go to the
// CTA's own tile (no inter-SM traffic; wrong results); 2 = cell update replaced by a sum (no SFU work; wrong
// results); 3 = product + timeline.
// SUBSEL: -1 = this warp serves both sub-tiles in turn; 0 / 1 = it serves only that sub-tile (split warp sets).
template <int NJ, int VARIANT, int SUBSEL>
__device__ __forceinline__ void epilogue_warp(const __half* __restrict__ gx, __half* __restrict__ y, int T, int N, int reverse,
                                              int n0, uint32_t rank, int blk, int quarter, int which, int col0,
                                              uint32_t tmem_base, uint32_t base, unsigned char* gbase, RecBars bars,
                                              int warp, int lane) {
    constexpr int NC = 8 * NJ;  // chunks of a sub-tile handled by this warp
    const int r = lane >> 2, q = lane & 3;
    const size_t gx_col = (size_t)rank * ROWS + (size_t)blk * 32 + r * 4;
    const int u0 = (int)rank * UPC + blk * 8;                                            // first unit of this block
    const int my_chunk = col0 + (lane & (NC - 1));                                       // chunk this lane writes to Y
    // destination inside a peer's h tile: k-chunk (u0/8), rows col0.. : NC*16 contiguous bytes
    const uint32_t dst_off = (uint32_t)(u0 >> 3) * (SN * 16) + (uint32_t)col0 * 16;
    // shared::cluster window of peer d relative to this CTA's (mapa is affine in the offset)
    uint32_t peer_shift[CS];
#pragma unroll
    for (int d = 0; d < CS; ++d) peer_shift[d] = mapa(base, VARIANT == 1 ? rank : (uint32_t)d) - base;
    float c_state[NS][NJ][2];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int j = 0; j < NJ; ++j) c_state[s][j][0] = c_state[s][j][1] = 0.f;

    // input pre-activations are prefetched one (step, sub-tile) item ahead: never on the recurrence's critical path
    constexpr int ITEMS = SUBSEL < 0 ? NS : 1;      // (step, sub-tile) items this warp walks per step
    const int wslot = warp & 7;                     // staging slot inside the (parity, sub) group
    auto load_gx = [&](int item, uint2 (&dst)[NJ][2]) {
        const int step = SUBSEL < 0 ? (item >> 1) : item, sub = SUBSEL < 0 ? (item & 1) : SUBSEL;
        const int t = reverse ? (T - 1 - step) : step;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = n0 + sub * SN + col0 + 8 * j + 2 * q + e;
                dst[j][e] = (n < N) ? __ldg(reinterpret_cast<const uint2*>(gx + ((size_t)t * N + n) * 4 * H + gx_col))
                                    : make_uint2(0, 0);
            }
    };
    // two items of lead (~2.8 us): under the HBM/L2 load of the GEMMs that run next to the recurrence a single item
    // of lead let the loads surface on the critical path (3.4 ms instead of 2.5 ms per layer-launch)
    uint2 g[NJ][2], gn[NJ][2], gnn[NJ][2];
    load_gx(0, g);
    if (T * ITEMS > 1) load_gx(1, gn);

    for (int step = 0; step < T; ++step) {
        const int t = reverse ? (T - 1 - step) : step;
        const int p = step & 1;
#pragma unroll
        for (int sub = 0; sub < NS; ++sub) {
            if (SUBSEL >= 0 && sub != SUBSEL) continue;
            const int item = SUBSEL < 0 ? step * NS + sub : step;
            // staging buffer (parity, sub): its last readers (bulk copies of step-2) are complete, see kernel comment
            const uint32_t stage_off = OFF_STAGE + (uint32_t)((p * NS + sub) * 8 + wslot) * STAGE_WARP;
            __half* stage = reinterpret_cast<__half*>(gbase + stage_off);
            if (item + 2 < T * ITEMS) load_gx(item + 2, gnn);
            mbar_wait(bars.dfull + 8 * sub, (uint32_t)(step & 1));
            const bool tl = VARIANT == 3 && sub == 0 && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 7);
            const int ts = step % TL_STEPS;
            if (tl && warp == 0) g_timeline[ts][2] = clock64();
            tc_fence_after();
            uint32_t a[4 * NJ], b[4 * NJ];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + COL_D + sub * 32 + which * 16 + col0;
            if (NJ == 2) {
                tc_ld_16x256b_x2(taddr, *reinterpret_cast<uint32_t(*)[8]>(a));                     // rows 0..15: gates i, f
                tc_ld_16x256b_x2(taddr + (16u << 16), *reinterpret_cast<uint32_t(*)[8]>(b));       // rows 16..31: gates g, o
            } else {
                tc_ld_16x256b_x1(taddr, *reinterpret_cast<uint32_t(*)[4]>(a));
                tc_ld_16x256b_x1(taddr + (16u << 16), *reinterpret_cast<uint32_t(*)[4]>(b));
            }
            tc_wait_ld();
            tc_fence_before();
            if (tl && warp == 0) g_timeline[ts][3] = clock64();
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const __half2 g01 = *reinterpret_cast<const __half2*>(&g[j][e].x);
                    const __half2 g23 = *reinterpret_cast<const __half2*>(&g[j][e].y);
                    const float ai = __uint_as_float(a[4 * j + e]) + __low2float(g01);
                    const float af = __uint_as_float(a[4 * j + 2 + e]) + __high2float(g01);
                    const float ag = __uint_as_float(b[4 * j + e]) + __low2float(g23);
                    const float ao = __uint_as_float(b[4 * j + 2 + e]) + __high2float(g23);
                    float c, h;
                    if (VARIANT == 2) {
                        c = 0.25f * (af + ai + ag) + 0.5f * c_state[sub][j][e];
                        h = 0.1f * (ao + c);
                    } else {
                        float si, sf, tg, so;
                        gate_activations(ai, af, ag, ao, si, sf, tg, so);
                        c = fmaf(sf, c_state[sub][j][e], si * tg);
                        h = so * tanh_f(c);
                    }
                    c_state[sub][j][e] = c;
                    stage[(8 * j + 2 * q + e) * 8 + r] = __float2half_rn(h);
                }
            fence_proxy_async_smem();   // staged block (generic stores) -> visible to the bulk-copy engine
            __syncwarp();
            if (tl && warp == 0) g_timeline[ts][4] = clock64();
            if (step + 1 < T && elect_one_sync()) {   // one lane: eight back-to-back bulk copies, one per peer
                const uint32_t dst = base + OFF_H + (uint32_t)(sub * 2 + (p ^ 1)) * HT + dst_off, src = base + stage_off;
                const uint32_t bar = bars.hfull + 8 * (sub * 2 + (p ^ 1));
#pragma unroll
                for (int d = 0; d < CS; ++d) bulk_copy_to_peer(dst + peer_shift[d], src, NC * 16, bar + peer_shift[d]);
            }
            if (lane < NC) {
                const uint4 chunk = reinterpret_cast<const uint4*>(stage)[lane];  // chunk col0+lane: its 8 units
                const int n = n0 + sub * SN + my_chunk;
                if (n < N) *reinterpret_cast<uint4*>(y + ((size_t)t * N + n) * H + u0) = chunk;
            }
            if (tl) g_timeline[ts][warp == 0 ? 5 : 7] = clock64();
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                g[j][0] = gn[j][0]; g[j][1] = gn[j][1];
                gn[j][0] = gnn[j][0]; gn[j][1] = gnn[j][1];
            }
            __syncwarp();
        }
    }
}

// Staging-buffer reuse: the block staged at step s (parity p, sub-tile u) is read asynchronously by 8 bulk copies.  It
// is overwritten at step s+2, after this CTA has seen its own h tile (u, s+2) complete, which needs every peer's
// epilogue of (u, s+1), which needs that peer's h tile (u, s+1) complete -- i.e. all copies of step s landed.
//
// Warp roles (9 warps).  TMEM lanes: accumulator 0 holds gate rows 0..127 (row blocks 0-3), accumulator 1 rows 64..191
// (blocks 2,3 again, then 4,5); a warp can only read the 32-lane quarter (warp % 4).  Work is spread so that every SM
// sub-partition gets 1.5 row blocks:   quarter 0         quarter 1         quarter 2          quarter 3
//   warps 0-3 (16 chunks/sub-tile)  block 0 (acc 0)    block 1 (acc 0)    block 4 (acc 1)     block 5 (acc 1)
//   warps 4-7 ( 8 chunks/sub-tile)  block 2 (acc 1) 0-7  block 3 (acc 1) 0-7  block 2 (acc 0) 8-15  block 3 (acc 0) 8-15
//   warp 8                          MMA issuer (+ TMEM allocation)
//
// SPLIT = 1: two sets of eight epilogue warps, one per sub-tile (17 warps), so the two sub-tiles' epilogues overlap in
// time instead of alternating on the same warps.
template <int VARIANT, int SPLIT>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(SPLIT ? THREADS_SPLIT : THREADS, 1)
lstm_rec_tc_kernel(const __half* __restrict__ gx, const __half* __restrict__ whh, __half* __restrict__ y, int T, int N,
                   int reverse) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    RecBars bars;
    bars.hfull = base + OFF_BARS;          // 4 barriers
    bars.dfull = bars.hfull + 8 * 4;       // 2 barriers
    const uint32_t tmem_slot = bars.dfull + 8 * 2;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int MMA_WARP = SPLIT ? 16 : 8;
    const uint32_t rank = cluster_ctarank();
    const int group = blockIdx.x / CS;
    const int n0 = group * NB;

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(bars.hfull + 8 * i, 1);
        for (int i = 0; i < 2; ++i) mbar_init(bars.dfull + 8 * i, 1);
        mbar_fence_init();
        // every fill of an h tile is SN*H*2 bytes of bulk-copy traffic from the 8 CTAs of the cluster
        for (int sub = 0; sub < NS; ++sub) {
            if (T > 1) mbar_expect_tx(bars.hfull + 8 * (sub * 2 + 1), HT);   // parity 1: filled during step 0
            if (T > 2) mbar_expect_tx(bars.hfull + 8 * (sub * 2 + 0), HT);   // parity 0: filled during step 1
        }
    }
    if (warp == MMA_WARP) tc_alloc(tmem_slot, TMEM_COLS);
    // h_{-1} = 0 (parity 0 tiles of both sub-tiles; zeroing everything is simplest)
    for (int i = tid; i < (int)(NS * 2 * HT / 16); i += (int)blockDim.x) reinterpret_cast<uint4*>(gbase + OFF_H)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + OFF_BARS + 8 * 6);

    // resident weights -> TMEM (lane = gate row, column c = fp16 pair (2c, 2c+1) of that row)
    if (warp < 4) {
        const __half* wsrc = whh + (size_t)rank * ROWS * H;
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
            const int row = blk * 64 + warp * 32 + lane;
            const uint4* src = reinterpret_cast<const uint4*>(wsrc + (size_t)row * H);
#pragma unroll 1
            for (int c = 0; c < H / 64; ++c) {
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint4 w = __ldg(src + c * 8 + i);
                    v[4 * i + 0] = w.x; v[4 * i + 1] = w.y; v[4 * i + 2] = w.z; v[4 * i + 3] = w.w;
                }
                tc_st_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (blk ? COL_A2 : COL_A1) + c * 32, v);
            }
        }
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_sync_all();  // every CTA's barriers are initialised before any peer's bulk copy can land

    if (warp == MMA_WARP) {
        // ===== MMA issuer: the whole warp walks the (step, sub-tile) items, one elected lane issues =====
        constexpr uint32_t idesc = tc_idesc_f16(128, SN);
        for (int step = 0; step < T; ++step) {
            const int p = step & 1;
#pragma unroll
            for (int sub = 0; sub < NS; ++sub) {
                const uint32_t hbar = bars.hfull + 8 * (sub * 2 + p);
                if (step > 0) mbar_wait(hbar, (uint32_t)((((step + 1) >> 1) - 1) & 1));
                if (elect_one_sync()) {
                    if (step > 0) {
                        if (step + 2 < T) mbar_expect_tx(hbar, HT);   // re-arm for the fill during step+1
                        fence_proxy_async_smem();
                    }
                    if (VARIANT == 3 && sub == 0 && blockIdx.x == 0) {
                        g_timeline[step % TL_STEPS][0] = clock64();
                        unsigned long long gt;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                        g_timeline[step % TL_STEPS][6] = (long long)gt;
                    }
                    tc_fence_after();
                    // B tile: [k-chunk][16 rows][16 B]; one K=16 step = two k-chunks = 512 B
                    const uint64_t bdesc0 = tc_smem_desc_noswz(base + OFF_H + (uint32_t)(sub * 2 + p) * HT, SN * 16, 128);
                    const uint32_t d1 = tmem_base + COL_D + sub * 32, d2 = d1 + 16;
#pragma unroll
                    for (int ks = 0; ks < H / 16; ++ks) {
                        const uint32_t acol = (uint32_t)ks * 8;
                        const uint32_t acc = ks != 0 ? 1u : 0u;
                        const uint64_t bdesc = bdesc0 + (uint64_t)(ks * (2 * SN * 16 / 16));
                        tc_mma_ts(d1, tmem_base + COL_A1 + acol, bdesc, idesc, acc);
                        tc_mma_ts(d2, tmem_base + COL_A2 + acol, bdesc, idesc, acc);
                    }
                    tc_commit(bars.dfull + 8 * sub);
                    if (VARIANT == 3 && sub == 0 && blockIdx.x == 0) g_timeline[step % TL_STEPS][1] = clock64();
                }
                __syncwarp();
            }
        }
    } else {
        const int quarter = warp & 3, ew = warp & 7;
        const int blk = ew < 4 ? (quarter < 2 ? quarter : quarter + 2)        // 0, 1, 4, 5
                               : 2 + (quarter & 1);                           // 2, 3, 2, 3
        const int which = ew < 4 ? (quarter < 2 ? 0 : 1) : (quarter < 2 ? 1 : 0);
        const int col0 = ew < 4 ? 0 : (quarter < 2 ? 0 : 8);
#define RUN_EPILOGUE(NJ_, SEL_)                                                                                          \
        epilogue_warp<NJ_, VARIANT, SEL_>(gx, y, T, N, reverse, n0, rank, blk, quarter, which, col0, tmem_base, base, gbase, \
                                          bars, warp, lane)
        if (!SPLIT) {
            if (ew < 4) RUN_EPILOGUE(2, -1); else RUN_EPILOGUE(1, -1);
        } else if (warp < 8) {
            if (ew < 4) RUN_EPILOGUE(2, 0); else RUN_EPILOGUE(1, 0);
        } else {
            if (ew < 4) RUN_EPILOGUE(2, 1); else RUN_EPILOGUE(1, 1);
        }
#undef RUN_EPILOGUE
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // nobody leaves while peers may still write into this CTA's shared memory
    if (warp == MMA_WARP) tc_dealloc(tmem_base, TMEM_COLS);
}

// ---- TMEM layout probe (debug / self-test) ----------------------------------------------------------
// out[0 .. 128*32): what each thread receives from tcgen05.ld.16x256b.x4 (lanes +0 and +16 of its quarter) after
//                   lane L / column c was filled with 100*L + c through tcgen05.st.32x32b;
// out[4096 .. 4096 + 128*32): D = A * B^T with A (128 x 16, fp16) taken from TMEM, B (32 x 16) from swizzled smem.
__global__ void __launch_bounds__(128, 1) tmem_probe_kernel(float* __restrict__ out) {
    __shared__ __align__(1024) unsigned char btile[NB * 128];
    __shared__ __align__(128) unsigned char btile_ns[4096];          // [k chunk][32 rows][16 B], no swizzle (+ slack)
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    if (warp == 0) tc_alloc(smem_u32(&slot), 256);
    for (int i = tid; i < NB * 128 / 4; i += 128) {
        reinterpret_cast<uint32_t*>(btile)[i] = 0;
        reinterpret_cast<uint32_t*>(btile_ns)[i] = 0;
    }
    __syncthreads();
    // B[n][k] = ((n + k) % 5) * 0.5, K-major SW128 rows (only the first 32 B of each 128-B row are used)
    if (tid < NB) {
        __half row[16];
        for (int k = 0; k < 16; ++k) row[k] = __float2half(((tid + k) % 5) * 0.5f);
        for (int c = 0; c < 2; ++c) {
            *reinterpret_cast<uint4*>(btile + sw128_offset(tid, c)) = *reinterpret_cast<uint4*>(row + 8 * c);
            *reinterpret_cast<uint4*>(btile_ns + c * (NB * 16) + tid * 16) = *reinterpret_cast<uint4*>(row + 8 * c);
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    {
        uint32_t v[32];
        for (int c = 0; c < 32; ++c) v[c] = __float_as_uint((float)(100 * tid + c));
        tc_st_32x32b_x32(tb + ((uint32_t)(warp * 32) << 16) + 0, v);
        // A[i][k] = ((i % 7) + k) * 0.25 packed as fp16 pairs in columns 32..39
        uint32_t w[32];
        for (int c = 0; c < 32; ++c) {
            __half2 h2 = __floats2half2_rn(((tid % 7) + 2 * c) * 0.25f, ((tid % 7) + 2 * c + 1) * 0.25f);
            w[c] = *reinterpret_cast<uint32_t*>(&h2);
        }
        tc_st_32x32b_x32(tb + ((uint32_t)(warp * 32) << 16) + 32, w);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        tc_mma_ts(tb + 64, tb + 32, tc_smem_desc_sw128(smem_u32(btile)), tc_idesc_f16(128, NB), 0u);
        // the same product with B in the un-swizzled core-matrix layout: (lbo, sbo) = (K step, row-group step) and swapped
        tc_mma_ts(tb + 96, tb + 32, tc_smem_desc_noswz(smem_u32(btile_ns), NB * 16, 128), tc_idesc_f16(128, NB), 0u);
        tc_mma_ts(tb + 128, tb + 32, tc_smem_desc_noswz(smem_u32(btile_ns), 128, NB * 16), tc_idesc_f16(128, NB), 0u);
        tc_commit(smem_u32(&bar));
    }
    {
        uint32_t a[16], b[16];
        tc_ld_16x256b_x4(tb + ((uint32_t)(warp * 32) << 16) + 0, a);
        tc_ld_16x256b_x4(tb + ((uint32_t)(warp * 32 + 16) << 16) + 0, b);
        tc_wait_ld();
        for (int i = 0; i < 16; ++i) {
            out[tid * 32 + i] = __uint_as_float(a[i]);
            out[tid * 32 + 16 + i] = __uint_as_float(b[i]);
        }
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    {
        uint32_t d[32];
        tc_ld_32x32b_x32(tb + ((uint32_t)(warp * 32) << 16) + 64, d);
        tc_wait_ld();
        for (int c = 0; c < 32; ++c) out[4096 + tid * 32 + c] = __uint_as_float(d[c]);
        tc_ld_32x32b_x32(tb + ((uint32_t)(warp * 32) << 16) + 96, d);
        tc_wait_ld();
        for (int c = 0; c < 32; ++c) out[8192 + tid * 32 + c] = __uint_as_float(d[c]);
        tc_ld_32x32b_x32(tb + ((uint32_t)(warp * 32) << 16) + 128, d);
        tc_wait_ld();
        for (int c = 0; c < 32; ++c) out[12288 + tid * 32 + c] = __uint_as_float(d[c]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tb, 256);
}

}  // namespace

bool lstm_rec_tc_supported(int hidden) { return hidden == H; }

int launch_lstm_rec_tc(const __half* gx, const __half* whh, __half* y, int T, int N, int hidden, int reverse,
                       cudaStream_t stream) {
    B200_REQUIRE(hidden == H, "lstm_rec_tc: hidden size %d is not supported (384)", hidden);
    const int groups = (N + NB - 1) / NB;
    const char* dbg = getenv("B200_LSTM_DEBUG");
    const int variant = dbg ? atoi(dbg) : 0;
    // experiments: B200_LSTM_SMEM_KB shrinks the shared-memory request (the split kernel cannot be co-resident with itself
    // anyway: 544 threads x 85 registers), which lets small CTAs of other kernels share the SM
    const char* skb = getenv("B200_LSTM_SMEM_KB");
    const uint32_t smem_bytes = skb && (uint32_t)atoi(skb) * 1024u >= SMEM_USED ? (uint32_t)atoi(skb) * 1024u : SMEM_BYTES;
    const char* sp = getenv("B200_LSTM_SPLIT");
    const bool split = sp ? atoi(sp) != 0 : true;   // measured: 2412 vs 2811 cycles per step (profiles/r01_lstm_split.md)
#define LAUNCH_VARIANT(v, s)                                                                                          \
    do {                                                                                                              \
        B200_CHECK_CUDA(cudaFuncSetAttribute(lstm_rec_tc_kernel<v, s>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                             (int)smem_bytes));                                                       \
        lstm_rec_tc_kernel<v, s><<<groups * CS, s ? THREADS_SPLIT : THREADS, smem_bytes, stream>>>(gx, whh, y, T, N,  \
                                                                                                   reverse);         \
    } while (0)
    if (variant == 3 && split) LAUNCH_VARIANT(3, 1);
    else if (variant == 3) LAUNCH_VARIANT(3, 0);
    else if (variant == 1) LAUNCH_VARIANT(1, 0);
    else if (variant == 2) LAUNCH_VARIANT(2, 0);
    else if (split) LAUNCH_VARIANT(0, 1);
    else LAUNCH_VARIANT(0, 0);
#undef LAUNCH_VARIANT
    B200_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// how many 8-CTA clusters of the recurrent kernel the device can hold at once (GPC packing decides)
int lstm_rec_tc_max_clusters() {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * 64);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaFuncSetAttribute(lstm_rec_tc_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess)
        return -1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, lstm_rec_tc_kernel<0, 0>, &cfg) != cudaSuccess) return -1;
    return n;
}

int copy_lstm_timeline(long long* host_out, int max_steps) {
    const int n = max_steps < TL_STEPS ? max_steps : TL_STEPS;
    B200_CHECK_CUDA(cudaDeviceSynchronize());
    B200_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(long long) * 8 * n));
    return n;
}

int launch_tmem_probe(float* out, cudaStream_t stream) {
    tmem_probe_kernel<<<1, 128, 0, stream>>>(out);
    B200_CHECK_CUDA(cudaGetLastError());
    return 0;
}
