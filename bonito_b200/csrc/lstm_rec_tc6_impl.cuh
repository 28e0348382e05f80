// Persistent recurrent part of one LSTM layer on the 5th-generation tensor cores (hac: H = 384), tile layout.
// Reference semantics: bonito/nn.py:353-415 (torch.nn.LSTM, gate order i,f,g,o, zero initial state, optional
// time reversal), the span `Model.use_koi` hands to koi.lstm (bonito/crf/model.py:240-246).
//
// Second-generation decomposition (the first one, lstm_rec_tc.cu, is kept for the generic [T][N][4H] layout):
//   * a cluster of CS = 6 CTAs owns one batch tile of NB = 48 chunks for all T steps; CTA `rank` owns hidden units
//     [64*rank, 64*rank+64) = 256 gate rows of W_hh = exactly two M=128 accumulators, no duplicated rows (the 8-CTA
//     split needed 192 rows per CTA = two overlapping 128-row blocks: a third of the tensor work was redundant).
//     22 such clusters fit on a B200 (GPC packing; 15 for clusters of 8), so the 11 tiles of a 512-chunk batch run in
//     ONE wave on 66 SMs and leave 82 SMs to the GEMMs / the decode of the neighbouring tiles and batches.
//   * W_hh slice lives in TENSOR MEMORY for the whole kernel as the UMMA A operand (2 x 192 columns of fp16 pairs).
//   * the tile is processed as THREE interleaved sub-tiles of 16 chunks, each with its own accumulator pair, its own
//     h tile (double buffered by step parity) and its own set of eight epilogue warps: while the epilogue warps of one
//     sub-tile run the cell update and the exchange, the tensor core works on the other two.  The per-step dependency
//     chain of a sub-tile (MMA -> tcgen05.ld -> cell update -> exchange -> next MMA) is ~2.2k cycles of which the tensor
//     pipe is busy ~0.55k, so three chains fill it.
//   * h_{t-1} of a sub-tile [16 chunks x 384] sits in shared memory as the UMMA B operand, K-major WITHOUT swizzle:
//     [48 k-chunks of 8 units][16 chunks][16 B], so the 8 units x 16 chunks one warp produces are 256 contiguous
//     bytes of every peer's tile.
//   * the input projection gx of this CTA for one (step, sub-tile) is ONE contiguous 8 KB block (the GEMM that
//     produces it writes the layout [tile][T][rank][48 chunks][256 columns]); the MMA warp streams it into a 4-deep
//     shared-memory ring per sub-tile with cp.async.bulk, three steps ahead: no global load, and no register prefetch
//     buffers, in the epilogue warps.
//   * per (step, sub-tile) one elected thread issues 2 x 24 tcgen05.mma (M=128, N=16, K=16) -- ONE issuing warp serves the
//     three sub-tiles in turn: giving each sub-tile its own issuing warp was measured slower (3474 vs 3089 cycles per
//     step), because the in-order issue is what keeps the three sub-tiles staggered: with three issuers they phase-lock,
//     their MMAs interleave in the tensor pipe and all 24 epilogue warps then run their cell updates at the same time;
//     eight epilogue warps pull
//     the gate pre-activations with tcgen05.ld.16x256b (the mma-accumulator fragment: with rows ordered
//     [8 units x (i,f,g,o)] one thread holds all four gates of a (unit, chunk)), add gx, update (c, h) in registers,
//     stage the new h block in shared memory and write it to Y[t].
//   * h all-gather.  Every CTA needs the whole h_t of a sub-tile (12 KB) every step.  Pushing the blocks peer by peer
//     through distributed shared memory (one cp.async.bulk shared::cta -> shared::cluster per peer, EXCH = 0) makes
//     every SM send AND receive 30 KB per step over its DSMEM port: 3480 cycles per step measured for this kernel with
//     it (2244 for the 8-CTA one: both ~18 B/clk per SM, in + out) -- the exchange bandwidth, not the tensor core, set
//     the step time.  EXCH = 1 (default) goes through L2: the warp writes its 256-byte block to a staging buffer in
//     global memory (which stays in L2), fences (fence.proxy.async.global) and issues ONE multicast bulk copy
//     (cp.async.bulk ... global -> shared::cluster, .multicast::cluster) that lands the block in the h tile of all six CTAs
//     and completes 256 bytes on each CTA's mbarrier: 24 TMA operations per CTA and step instead of 144 and no DSMEM
//     traffic.  Exchange skeleton alone (scripts/exchange_bench.py): 840 vs 1830 cycles per step.  (A multicast TENSOR
//     load straight out of Y, box = 8 units x 16 chunks, needs no staging buffer but measured 10700 cycles per step: its
//     16-byte box rows are one L2 request each.)
//
// Operands: whh [4H][H] rows permuted [unit/8][gate][unit%8] (rank r owns rows 256r..256r+255);
//           gx  [tile][T][6][48][256]  columns of rank r = [unit/8 - 8r][unit%8][gate];   y [tile][T][48][H].
// This file is the kernel BODY: lstm_rec_tc6.cu includes it once per tile shape (LSTM6_NS sub-tiles of LSTM6_SN chunks,
// ring depth LSTM6_GXD) inside its own namespace LSTM6_NAMESPACE.

namespace LSTM6_NAMESPACE {

constexpr int NS = LSTM6_NS;       // interleaved sub-tiles
constexpr int SN = LSTM6_SN;       // chunks per sub-tile = N of one MMA (16 or 32)
constexpr int NJ = SN / 8;         // 8-chunk column groups per epilogue warp (the warp handles all SN chunks of its row block)
constexpr int NB = NS * SN;        // chunks per cluster (tile)
constexpr int H = 384;
constexpr int CS = 6;
constexpr int UPC = H / CS;        // 64 units per CTA
constexpr int ROWS = 4 * UPC;      // 256 gate rows per CTA
constexpr int EW = 8;              // epilogue warps per sub-tile (one 32-row block each)
constexpr int MMA_WARP = NS * EW;  // warp 24: MMA issuer + gx producer + TMEM allocation
constexpr int THREADS = (MMA_WARP + 1) * 32;      // 800
constexpr int GXD = LSTM6_GXD;     // depth of the gx ring (steps)
constexpr uint32_t HT = (H / 8) * SN * 16;        // one h tile: 48 k-chunks x 16 chunks x 16 B = 12288 B
constexpr uint32_t GXS = SN * ROWS * 2;           // gx of one (step, sub-tile) for this CTA: 8192 B
constexpr uint32_t COL_A1 = 0, COL_A2 = 192, COL_D = 384, TMEM_COLS = 512;   // D: [sub][acc 0|1] x SN columns
static_assert(COL_D + NS * 2 * SN <= TMEM_COLS, "accumulators do not fit next to the resident weights");
constexpr uint32_t STAGE_WARP = SN * 16;          // 256 B per (parity, sub, warp)
constexpr uint32_t OFF_H = 0;                                        // [sub][parity] h tiles
constexpr uint32_t OFF_STAGE = NS * 2 * HT;                          // [parity][sub][warp]
constexpr uint32_t OFF_GX = OFF_STAGE + 2 * NS * EW * STAGE_WARP;    // [sub][slot]
constexpr uint32_t OFF_BARS = OFF_GX + NS * GXD * GXS;
constexpr uint32_t N_BARS = NS * 2 + NS + NS * GXD;                  // hfull, dfull, gxfull
constexpr uint32_t SMEM_BYTES = OFF_BARS + 8 * N_BARS + 64 + 1024;   // 182 / 211 KB: one CTA per SM (it owns all 512 TMEM columns)
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// timeline of CTA 0 (VARIANT 3), sub-tile 0: per step, SM-clock stamps of
//   [0] h tile complete (MMA thread)   [1] MMAs issued + committed   [2] accumulator ready (epilogue warp 0)
//   [3] TMEM loaded   [4] cell update done + staged   [5] h block sent (warp 0)   [6] %globaltimer (ns) at [0]
//   [7] the MMA warp starts waiting for the h tile (its work of the previous step is issued)
constexpr int TL_STEPS = 256;
__device__ long long g_timeline6[TL_STEPS][8];

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

__device__ __forceinline__ void bulk_load_global(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
// multicast bulk copy global (L2) -> the same CTA-relative shared-memory offset of every CTA in `mask`; each destination
// CTA's mbarrier (same CTA-relative offset) receives `bytes` of complete_tx
__device__ __forceinline__ void bulk_multicast(uint32_t dst, const void* gsrc, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n" ::
            "r"(dst), "l"(gsrc), "r"(bytes), "r"(bar), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;\n" ::: "memory"); }
__device__ __forceinline__ uint2 lds_v2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

struct Bars {
    uint32_t base;
    __device__ __forceinline__ uint32_t hfull(int sub, int parity) const { return base + 8u * (uint32_t)(sub * 2 + parity); }
    __device__ __forceinline__ uint32_t dfull(int sub) const { return base + 8u * (uint32_t)(NS * 2 + sub); }
    __device__ __forceinline__ uint32_t gxfull(int sub, int slot) const { return base + 8u * (uint32_t)(NS * 3 + sub * GXD + slot); }
};

// One epilogue warp of sub-tile `sub`: row block `blk` = 8 hidden units x 4 gates (32 TMEM lanes at lane quarter
// blk % 4 of accumulator blk / 4), all 16 chunks of the sub-tile.
// VARIANT (B200_LSTM_DEBUG): 0 = product; 3 = product + timeline.
template <int VARIANT, int EXCH>
__device__ __forceinline__ void epilogue_warp(__half* __restrict__ y, unsigned char* __restrict__ hx,
                                              const __half* __restrict__ gx_sub, int T, int nb,
                                              int reverse, uint32_t rank, int sub, int ew, uint32_t tmem_base,
                                              uint32_t base, unsigned char* gbase, Bars bars, int lane, int ablate) {
    const int r = lane >> 2, q = lane & 3;
    const int quarter = ew & 3, which = ew >> 2, blk = which * 4 + quarter;
    const int u0 = (int)rank * UPC + blk * 8;                                   // first unit of this block
    // destination inside a peer's h tile: k-chunk (u0 / 8): 256 contiguous bytes
    const uint32_t dst_off = (uint32_t)(u0 >> 3) * (SN * 16);
    // shared::cluster window of peer d relative to this CTA's (mapa is affine in the offset); own rank last
    uint32_t peer_shift[EXCH == 0 ? CS : 1];
    if (EXCH == 0) {
#pragma unroll
        for (int d = 0; d < CS; ++d) peer_shift[d] = mapa(base, (rank + 1u + (uint32_t)d) % CS) - base;
    }
    float c_state[NJ][2];
#pragma unroll
    for (int j = 0; j < NJ; ++j) c_state[j][0] = c_state[j][1] = 0.f;
    // this lane's gx values inside a ring slot: chunk 8j + 2q + e, columns blk*32 + r*4 .. +3 (gates i,f,g,o of unit r)
    const uint32_t gx_lane = base + OFF_GX + (uint32_t)(sub * GXD) * GXS + (uint32_t)(2 * q) * (ROWS * 2) + (uint32_t)(blk * 32 + r * 4) * 2;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + COL_D + (uint32_t)(sub * 2 * SN + which * SN);
    const bool y_ok = lane < SN && sub * SN + lane < nb;
    __half* y_lane = y + (size_t)(sub * SN + (lane & (SN - 1))) * H + u0;

    for (int step = 0; step < T; ++step) {
        const int t = reverse ? (T - 1 - step) : step;
        const int p = step & 1;
        const int slot = step % GXD;
        // staging buffer (parity, sub): its last readers (bulk copies of step-2) are complete, see kernel comment
        const uint32_t stage_off = OFF_STAGE + (uint32_t)((p * NS + sub) * EW + ew) * STAGE_WARP;
        __half* stage = reinterpret_cast<__half*>(gbase + stage_off);
        // input projection of this step from the ring (landed ~3 steps ago: never on the recurrence's critical path)
        if (!(ablate & 1)) mbar_wait(bars.gxfull(sub, slot), (uint32_t)((step / GXD) & 1));
        uint2 g[NJ][2];
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) g[j][e] = lds_v2(gx_lane + (uint32_t)slot * GXS + (uint32_t)(8 * j + e) * (ROWS * 2));
        mbar_wait(bars.dfull(sub), (uint32_t)(step & 1));
        const bool tl = VARIANT == 3 && sub == 0 && blockIdx.x == 0 && lane == 0 && (ew == 0 || ew == 7);
        const int ts = step % TL_STEPS;
        if (tl && ew == 0) g_timeline6[ts][2] = clock64();
        tc_fence_after();
        uint32_t a[4 * NJ], b[4 * NJ];
        if constexpr (NJ == 2) {
            tc_ld_16x256b_x2(taddr, *reinterpret_cast<uint32_t(*)[8]>(a));                  // rows 0..15 of the block: gates i, f
            tc_ld_16x256b_x2(taddr + (16u << 16), *reinterpret_cast<uint32_t(*)[8]>(b));    // rows 16..31: gates g, o
        } else {
            tc_ld_16x256b_x4(taddr, *reinterpret_cast<uint32_t(*)[16]>(a));
            tc_ld_16x256b_x4(taddr + (16u << 16), *reinterpret_cast<uint32_t(*)[16]>(b));
        }
        tc_wait_ld();
        tc_fence_before();
        if (tl && ew == 0) g_timeline6[ts][3] = clock64();
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
                float si, sf, tg, so, c, hval;
                if (ablate & 4) {      // timing experiment: no SFU work (wrong results)
                    c = 0.25f * (af + ai + ag) + 0.5f * c_state[j][e];
                    hval = 0.1f * (ao + c);
                } else {
                    gate_activations(ai, af, ag, ao, si, sf, tg, so);
                    c = fmaf(sf, c_state[j][e], si * tg);
                    hval = so * tanh_f(c);
                }
                c_state[j][e] = c;
                stage[(8 * j + 2 * q + e) * 8 + r] = __float2half_rn(hval);
            }
        if (EXCH == 0) {
            fence_proxy_async_smem();   // staged block (generic stores) -> visible to the bulk-copy engine
            __syncwarp();
            if (tl && ew == 0) g_timeline6[ts][4] = clock64();
            if (step + 1 < T && elect_one_sync()) {   // one lane: six back-to-back bulk copies, one per peer
                const uint32_t dst = base + OFF_H + (uint32_t)(sub * 2 + (p ^ 1)) * HT + dst_off, src = base + stage_off;
                const uint32_t bar = bars.hfull(sub, p ^ 1);
#pragma unroll
                for (int d = 0; d < CS; ++d) bulk_copy_to_peer(dst + peer_shift[d], src, STAGE_WARP, bar + peer_shift[d]);
            }
            if (y_ok) {   // chunk `lane` of the sub-tile: its 8 units
                const uint4 chunk = reinterpret_cast<const uint4*>(stage)[lane];
                *reinterpret_cast<uint4*>(y_lane + (size_t)t * (NB * H)) = chunk;
            }
        } else {
            __syncwarp();
            if (tl && ew == 0) g_timeline6[ts][4] = clock64();
            // staging block of this warp for the tile (sub, parity p^1): [k-chunk u0/8][16 chunks][16 B], the h-tile layout
            unsigned char* g = hx + (size_t)((p ^ 1) * NS + sub) * HT + dst_off;
            uint4 chunk = make_uint4(0, 0, 0, 0);
            if (lane < SN) {   // chunk `lane` of the sub-tile: its 8 units -> the staging block first (it is on the critical path)
                chunk = reinterpret_cast<const uint4*>(stage)[lane];
                if (step + 1 < T) reinterpret_cast<uint4*>(g)[lane] = chunk;
            }
            if (step + 1 < T) {
                fence_proxy_async_global();   // the block just written (generic proxy) -> visible to the TMA (async proxy)
                __syncwarp();
                if (elect_one_sync())
                    bulk_multicast(base + OFF_H + (uint32_t)(sub * 2 + (p ^ 1)) * HT + dst_off, g, STAGE_WARP,
                                   bars.hfull(sub, p ^ 1), (uint16_t)((1u << CS) - 1u));
            }
            if (y_ok && !(ablate & 2)) *reinterpret_cast<uint4*>(y_lane + (size_t)t * (NB * H)) = chunk;   // -> Y[t], off the critical path
        }
        if (tl && ew == 0) g_timeline6[ts][5] = clock64();
        // gx refill: this warp is past the accumulator barrier of `step`, so the h tile (sub, step) was complete, so every
        // epilogue warp of the sub-tile has sent -- and therefore consumed its gx of -- step-1: that ring slot is free
        if (ew == 0 && step + GXD - 1 < T && !(ablate & 1) && elect_one_sync()) {
            const int s2 = step + GXD - 1, t2 = reverse ? (T - 1 - s2) : s2, slot2 = s2 % GXD;
            const uint32_t bar = bars.gxfull(sub, slot2);
            mbar_expect_tx(bar, GXS);
            bulk_load_global(base + OFF_GX + (uint32_t)(sub * GXD + slot2) * GXS, gx_sub + (size_t)t2 * (CS * NB * ROWS), GXS, bar);
        }
        __syncwarp();
    }
}

// Staging-buffer reuse: the block staged at step s (parity p, sub-tile u) is read asynchronously by 6 bulk copies.  It
// is overwritten at step s+2, after this CTA has seen its own h tile (u, s+2) complete, which needs every peer's
// epilogue of (u, s+1), which needs that peer's h tile (u, s+1) complete -- i.e. all copies of step s landed.
// gx ring reuse: epilogue warp 0 of sub-tile u refills the slot of step s-1 (with step s+GXD-1) during step s, after the
// accumulator barrier of (u, s): the MMAs of (u, s) were issued after the h tile (u, s) was complete, which needs this
// CTA's own epilogue warps of (u, s-1) to have sent, i.e. to have consumed gx (u, s-1).
// TMEM accumulator reuse: the MMAs of (u, s+1) are issued after the h tile (u, s+1) is complete, i.e. after every
// epilogue warp of (u, s) has drained its accumulator block (tcgen05.wait::ld precedes the send).
template <int VARIANT, int EXCH>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(THREADS, 1)
lstm_rec_tc6_kernel(const __half* __restrict__ gx, const __half* __restrict__ whh, __half* __restrict__ y,
                    unsigned char* __restrict__ hx, int T, int N, int reverse, int ablate) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    Bars bars;
    bars.base = base + OFF_BARS;
    const uint32_t tmem_slot = bars.base + 8 * N_BARS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    const int tile = blockIdx.x / CS;
    const int nb = min(NB, N - tile * NB);          // valid chunks of this tile
    const int nsub = (nb + SN - 1) / SN;            // active sub-tiles (the same in every CTA of the cluster)
    gx += (size_t)tile * T * (CS * NB * ROWS);
    y += (size_t)tile * T * (NB * H);
    hx += (size_t)tile * (2 * NS * HT);            // this tile's exchange staging: [parity][sub][h tile]

    if (tid == 0) {
        for (uint32_t i = 0; i < N_BARS; ++i) mbar_init(bars.base + 8 * i, 1);
        mbar_fence_init();
        // every fill of an h tile is SN*H*2 bytes of bulk-copy traffic from the 6 CTAs of the cluster
        for (int sub = 0; sub < nsub; ++sub) {
            if (T > 1) mbar_expect_tx(bars.hfull(sub, 1), HT);   // parity 1: filled during step 0
            if (T > 2) mbar_expect_tx(bars.hfull(sub, 0), HT);   // parity 0: filled during step 1
        }
    }
    if (warp == MMA_WARP) tc_alloc(tmem_slot, TMEM_COLS);
    // h_{-1} = 0 (parity 0 tiles; zeroing everything is simplest)
    for (int i = tid; i < (int)(NS * 2 * HT / 16); i += THREADS) reinterpret_cast<uint4*>(gbase + OFF_H)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + OFF_BARS + 8 * N_BARS);

    // resident weights -> TMEM (lane = gate row, column c = fp16 pair (2c, 2c+1) of that row)
    if (warp < 4) {
        const __half* wsrc = whh + (size_t)rank * ROWS * H;
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
            const int row = blk * 128 + warp * 32 + lane;
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
        // ===== MMA issuer + gx producer: the whole warp walks the (step, sub-tile) items, one elected lane issues =====
        constexpr uint32_t idesc = tc_idesc_f16(128, SN);
        const __half* gx_cta = gx + (size_t)rank * (NB * ROWS);
        auto load_gx = [&](int step, int sub) {     // gx of (step, sub) -> ring slot step % GXD
            const int t = reverse ? (T - 1 - step) : step;
            const int slot = step % GXD;
            const uint32_t bar = bars.gxfull(sub, slot);
            mbar_expect_tx(bar, GXS);
            bulk_load_global(base + OFF_GX + (uint32_t)(sub * GXD + slot) * GXS,
                             gx_cta + (size_t)t * (CS * NB * ROWS) + (size_t)sub * (SN * ROWS), GXS, bar);
        };
        if (elect_one_sync()) {
            for (int s = 0; s < GXD - 1 && s < T && !(ablate & 1); ++s)
                for (int sub = 0; sub < nsub; ++sub) load_gx(s, sub);
        }
        __syncwarp();
        for (int step = 0; step < T; ++step) {
            const int p = step & 1;
#pragma unroll
            for (int sub = 0; sub < NS; ++sub) {
                if (sub >= nsub) break;
                const uint32_t hbar = bars.hfull(sub, p);
                if (VARIANT == 3 && sub == 0 && blockIdx.x == 0 && lane == 0) g_timeline6[step % TL_STEPS][7] = clock64();
                if (step > 0) mbar_wait(hbar, (uint32_t)((((step + 1) >> 1) - 1) & 1));
                if (elect_one_sync()) {
                    if (step > 0) {
                        if (step + 2 < T) mbar_expect_tx(hbar, HT);   // re-arm for the fill during step+1
                        // (no proxy fence: the h tile was written by bulk copies and is read by the tensor core, both async proxy)
                    }
                    if (VARIANT == 3 && sub == 0 && blockIdx.x == 0) {
                        g_timeline6[step % TL_STEPS][0] = clock64();
                        unsigned long long gt;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                        g_timeline6[step % TL_STEPS][6] = (long long)gt;
                    }
                    tc_fence_after();
                    // B tile: [k-chunk][16 rows][16 B]; one K=16 step = two k-chunks = 512 B
                    const uint64_t bdesc0 = tc_smem_desc_noswz(base + OFF_H + (uint32_t)(sub * 2 + p) * HT, SN * 16, 128);
                    const uint32_t d1 = tmem_base + COL_D + sub * 2 * SN, d2 = d1 + SN;
#pragma unroll
                    for (int ks = 0; ks < H / 16; ++ks) {
                        const uint32_t acol = (uint32_t)ks * 8;
                        const uint32_t acc = ks != 0 ? 1u : 0u;
                        const uint64_t bdesc = bdesc0 + (uint64_t)(ks * (2 * SN * 16 / 16));
                        tc_mma_ts(d1, tmem_base + COL_A1 + acol, bdesc, idesc, acc);
                        tc_mma_ts(d2, tmem_base + COL_A2 + acol, bdesc, idesc, acc);
                    }
                    tc_commit(bars.dfull(sub));
                    if (VARIANT == 3 && sub == 0 && blockIdx.x == 0) g_timeline6[step % TL_STEPS][1] = clock64();
                    // the gx refill for step + GXD - 1 is issued by epilogue warp 0 of the sub-tile (keeps ~100 cycles per
                    // sub-tile out of this loop: the issuing warp's ~3 x 900 cycles per step bounded the step time)
                }
                __syncwarp();
            }
        }
    } else {
        const int sub = warp / EW, ew = warp % EW;
        if (sub < nsub)
            epilogue_warp<VARIANT, EXCH>(y, hx, gx + (size_t)rank * (NB * ROWS) + (size_t)sub * (SN * ROWS), T, nb, reverse, rank, sub,
                                         ew, tmem_base, base, gbase, bars, lane, ablate);
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // nobody leaves while peers may still write into this CTA's shared memory
    if (warp == MMA_WARP) tc_dealloc(tmem_base, TMEM_COLS);
}


size_t workspace_bytes(int N) { return (size_t)((N + NB - 1) / NB) * (2 * NS * HT); }

// gx [tiles][T][6][48][256], y [tiles][T][48][H]; tiles = ceil(N / 48), the last one may be partial;
// workspace: lstm_rec_tile_workspace_bytes(N) bytes of exchange staging (contents irrelevant)
int launch(const __half* gx, const __half* whh, __half* y, void* workspace, int T, int N, int hidden,
                        int reverse, cudaStream_t stream) {
    B200_REQUIRE(hidden == H, "lstm_rec_tile: hidden size %d is not supported (384)", hidden);
    B200_REQUIRE(((uintptr_t)gx % 16) == 0 && ((uintptr_t)y % 16) == 0 && ((uintptr_t)whh % 16) == 0 &&
                     ((uintptr_t)workspace % 16) == 0,
                 "lstm_rec_tile: operands must be 16-byte aligned");
    const int tiles = (N + NB - 1) / NB;
    const char* dbg = getenv("B200_LSTM_DEBUG");
    const int variant = dbg ? atoi(dbg) : 0;
    // default: multicast bulk copies out of the L2 staging buffer; "dsmem": peer-to-peer bulk copies (cross-check)
    const char* ex = getenv("B200_LSTM_EXCH");
    const bool dsmem = ex && ex[0] == 'd';
    // B200_LSTM_ABLATE (timing experiments, wrong results): 1 = no gx traffic, 2 = no Y stores, 4 = no SFU work in the cell update
    const char* ab = getenv("B200_LSTM_ABLATE");
    const int ablate = ab ? atoi(ab) : 0;
#define LAUNCH6(v, e)                                                                                                   \
    do {                                                                                                                \
        B200_CHECK_CUDA(cudaFuncSetAttribute(lstm_rec_tc6_kernel<v, e>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                             (int)SMEM_BYTES));                                                         \
        lstm_rec_tc6_kernel<v, e><<<tiles * CS, THREADS, SMEM_BYTES, stream>>>(gx, whh, y, (unsigned char*)workspace, T, \
                                                                               N, reverse, ablate);                     \
    } while (0)
    if (variant == 3 && dsmem) LAUNCH6(3, 0);
    else if (variant == 3) LAUNCH6(3, 1);
    else if (dsmem) LAUNCH6(0, 0);
    else LAUNCH6(0, 1);
#undef LAUNCH6
    B200_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int copy_timeline(long long* host_out, int max_steps) {
    const int n = max_steps < TL_STEPS ? max_steps : TL_STEPS;
    B200_CHECK_CUDA(cudaDeviceSynchronize());
    B200_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_timeline6, sizeof(long long) * 8 * n));
    return n;
}

}  // namespace LSTM6_NAMESPACE
