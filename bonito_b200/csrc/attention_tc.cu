// Windowed softmax attention on the 5th-generation tensor cores (sup v5: head_dim 64, window 127 / 128).
// Reference semantics: bonito/transformer/model.py:42-79 -- flash_attn_qkvpacked_func(window_size=(wl, wr)), non-causal,
// softmax scale 1/sqrt(head_dim); the rotary embedding has been applied to q and k in place before (rotary_kernel).
//
// Persistent kernel, one CTA per SM; a work item is one (chunk, head): the CTA walks its 128-query tiles in order.  With
// wl, wr <= 128 the queries of tile qt see at most the three 128-key tiles qt-1, qt, qt+1, so
//   * K and V tiles travel through 4-slot shared-memory rings (TMA, 3-D tensor map over qkv [N][T][3*heads*64],
//     SWIZZLE_128B, out-of-range rows zero-filled): every key tile is loaded ONCE per (chunk, head) and used by three
//     query tiles; the tile needed next is always in flight while the current one is processed; Q is double buffered;
//   * S_j = Q K_j^T: 4 tcgen05.mma (M=128, N=128, K=16) per key tile, fp32 accumulators in TMEM columns [128j, 128j+128);
//   * eight softmax warps, two per TMEM lane quarter (tcgen05.ld 32x32b: thread i of a warp owns lane 32 (warp % 4) + i = one
//     query row; the two warps of a row split each key tile's 128 columns in halves and exchange their row maxima through
//     shared memory): each key tile is normalised on its OWN row maximum m_j -- the scores live in registers between the
//     maximum and the exponentials, so S is read from TMEM exactly once -- and P_j = 2^((s - m_j) * scale) is written back
//     as fp16 pairs into the first 64 columns of S_j (tcgen05.st), i.e. as a TMEM-resident A operand.  32-column pieces
//     that a whole warp sees unmasked skip the masking pass; pieces masked for the whole warp are not even read;
//   * O_j = P_j V_j: 8 tcgen05.mma (M=128, N=64, K=16) per key tile with A from TMEM and B = V_j straight from its TMA
//     layout (rows = keys = K, 128-byte rows of 64 head dims: the MN-major SWIZZLE_128B operand, "transpose B" bit of the
//     instruction descriptor), accumulated into the LAST 64 columns of S_j; PV_j runs while the softmax warps work on j+1;
//   * epilogue: O = sum_j O_j 2^((m_j - M) scale) / sum_j l_j 2^((m_j - M) scale), M = max_j m_j -- three independent
//     partial softmaxes combined per row, no running-maximum rescaling of accumulators in flight.
// (First version: one CTA per query tile, nothing prefetched: 41 ms per sup step against 36 ms for the mma.sync kernel --
// every CTA paid the full load latency and a TMEM allocation for 128 queries.)
#include <cuda.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace {

constexpr int BQ = 128, BKV = 128, HD = 64, NKT = 3, RING = 4;
constexpr int SM_WARPS = 8;                        // softmax warps: two per TMEM lane quarter, each takes half of the columns
constexpr int THREADS = (SM_WARPS + 1) * 32;       // + 1 TMA / MMA warp
constexpr uint32_t TILE_BYTES = BQ * HD * 2;       // 16 KB: one 128 x 64 fp16 tile (128-byte rows)
constexpr uint32_t OFF_Q = 0, OFF_K = 2 * TILE_BYTES, OFF_V = OFF_K + RING * TILE_BYTES, OFF_BARS = OFF_V + RING * TILE_BYTES;
constexpr uint32_t OFF_XCH = OFF_BARS + 256;       // float [4][2][BQ]: per-row maxima / sums exchanged between the two warps of a row
constexpr uint32_t SMEM_BYTES = OFF_XCH + 4 * 2 * BQ * 4 + 1024;   // ~166 KB
constexpr uint32_t TMEM_COLS = 512;                // S_j | P_j | O_j share columns [128j, 128j+128), j < 3

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::
            "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// timeline of CTA 0 (B200_ATTN_DEBUG=1): per query tile, SM-clock stamps of
//   MMA thread: [0] Q ready + TMEM free  [1] QK MMAs issued  [2] PV MMAs issued (all p(j) seen)  [3] O committed + complete  [4] loads issued
//   softmax warp 0: [5] s(0) seen [6] S_0 loaded [7] max exchanged [8] P_0 stored + arrived [9] s(1) seen [10] P_1 arrived
//                   [11] P_2 arrived [12] o seen [13] O loaded [14] output stored
constexpr int ATL_TILES = 64;
__device__ long long g_attn_timeline[ATL_TILES][16];

struct AttnBars {
    uint32_t base;
    __device__ __forceinline__ uint32_t qfull(int b) const { return base + 8u * (uint32_t)b; }              // 2
    __device__ __forceinline__ uint32_t kvfull(int slot) const { return base + 16u + 8u * (uint32_t)slot; }   // RING (K and V of a key tile)
    __device__ __forceinline__ uint32_t s(int j) const { return base + 48u + 8u * (uint32_t)j; }              // 3
    __device__ __forceinline__ uint32_t p(int j) const { return base + 72u + 8u * (uint32_t)j; }              // 3
    __device__ __forceinline__ uint32_t o() const { return base + 96u; }
    __device__ __forceinline__ uint32_t tfree() const { return base + 104u; }
};

__global__ void __launch_bounds__(THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, __half* __restrict__ out, int N, int T, int NH, int wl,
                    int wr, float scale_log2e, int debug) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024-B alignment
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    AttnBars bars;
    bars.base = base + OFF_BARS;
    const uint32_t tmem_slot = bars.base + 112;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nqt = (T + BQ - 1) / BQ;             // query tiles per (chunk, head); key tiles -1 .. nqt are loaded
    const int items = N * NH;
    const bool tl = debug != 0 && blockIdx.x == 0;
#define ATL(it_, k_) do { if (tl && (it_) < ATL_TILES) g_attn_timeline[(it_)][(k_)] = clock64(); } while (0)

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_qkv));
        for (int b = 0; b < 2; ++b) mbar_init(bars.qfull(b), 1);
        for (int sl = 0; sl < RING; ++sl) mbar_init(bars.kvfull(sl), 1);
        for (int j = 0; j < NKT; ++j) {
            mbar_init(bars.s(j), 1);
            mbar_init(bars.p(j), SM_WARPS);        // one arrive per softmax warp
        }
        mbar_init(bars.o(), 1);
        mbar_init(bars.tfree(), SM_WARPS);
        mbar_fence_init();
    }
    if (warp == SM_WARPS) tc_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + OFF_BARS + 112);

    if (warp == SM_WARPS) {
        // ===== TMA producer + MMA issuer (one thread) =====
        if (elect_one_sync()) {
            constexpr uint32_t idesc_s = tc_idesc_f16(BQ, BKV);
            constexpr uint32_t idesc_o = tc_idesc_f16(BQ, HD) | (1u << 16);   // bit 16: B is MN-major
            // Flat streams over all work items of this CTA: query tiles `it` (Q buffer it & 1) and key-tile loads `kl`
            // (ring slot kl & 3).  Item i contributes nqt query tiles and nqt + 2 key tiles (kt = -1 .. nqt); query tile
            // qt of the item uses the key-tile loads first_kl + qt, +1, +2.
            int it = 0;                    // query tiles processed
            int kl_issued = 0;             // key tiles whose loads have been issued
            int q_issued = 0;              // query tiles whose Q load has been issued
            int kl_released = 0;           // key tiles no query tile will read again (their ring slots may be refilled)
            const int my_items = blockIdx.x < items ? (items - blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
            const int total_q = my_items * nqt, total_k = my_items * (nqt + 2);

            auto issue_loads = [&]() {
                // key tiles: up to RING in flight beyond the released ones
                while (kl_issued < total_k && kl_issued < kl_released + RING) {
                    const int li = kl_issued / (nqt + 2), kt = kl_issued % (nqt + 2) - 1;
                    const int item = blockIdx.x + li * (int)gridDim.x, n = item / NH, h = item % NH;
                    const int sl = kl_issued & (RING - 1);
                    mbar_expect_tx(bars.kvfull(sl), 2 * TILE_BYTES);
                    tma_load_3d(base + OFF_K + sl * TILE_BYTES, &map_qkv, bars.kvfull(sl), NH * HD + h * HD, kt * BKV, n);
                    tma_load_3d(base + OFF_V + sl * TILE_BYTES, &map_qkv, bars.kvfull(sl), 2 * NH * HD + h * HD, kt * BKV, n);
                    ++kl_issued;
                }
                // Q: the tile being processed and the next one
                while (q_issued < total_q && q_issued < it + 2) {
                    const int li = q_issued / nqt, qt = q_issued % nqt;
                    const int item = blockIdx.x + li * (int)gridDim.x, n = item / NH, h = item % NH;
                    mbar_expect_tx(bars.qfull(q_issued & 1), TILE_BYTES);
                    tma_load_3d(base + OFF_Q + (q_issued & 1) * TILE_BYTES, &map_qkv, bars.qfull(q_issued & 1), h * HD, qt * BQ, n);
                    ++q_issued;
                }
            };

            issue_loads();
            for (int li = 0; li < my_items; ++li) {
                const int first_kl = li * (nqt + 2);
                for (int qt = 0; qt < nqt; ++qt, ++it) {
                    // S_j = Q K_j^T  (both operands K-major, 128-byte swizzled rows of 64 head dims)
                    mbar_wait(bars.qfull(it & 1), (uint32_t)((it >> 1) & 1));
                    if (it > 0) mbar_wait(bars.tfree(), (uint32_t)((it - 1) & 1));    // the previous tile's O has been read
                    tc_fence_after();
                    ATL(it, 0);
                    const uint64_t qdesc = tc_smem_desc_sw128(base + OFF_Q + (it & 1) * TILE_BYTES);
#pragma unroll
                    for (int j = 0; j < NKT; ++j) {
                        const int kl = first_kl + qt + j, sl = kl & (RING - 1);
                        mbar_wait(bars.kvfull(sl), (uint32_t)((kl / RING) & 1));
                        tc_fence_after();
                        const uint64_t kdesc = tc_smem_desc_sw128(base + OFF_K + sl * TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k)
                            tc_mma_ss(tmem_base + (uint32_t)(j * 128), qdesc + 2u * k, kdesc + 2u * k, idesc_s, k != 0 ? 1u : 0u);
                        tc_commit(bars.s(j));
                    }
                    ATL(it, 1);
                    // O_j = P_j V_j  (A = P_j from TMEM: lane = query, column c = keys 2c, 2c+1; B = V_j MN-major: K = keys are
                    // the 128-byte rows, 8 keys per 1024-byte swizzle atom: one K = 16 step advances the descriptor by 2048 B)
#pragma unroll
                    for (int j = 0; j < NKT; ++j) {
                        const int kl = first_kl + qt + j, sl = kl & (RING - 1);
                        mbar_wait(bars.p(j), (uint32_t)(it & 1));
                        tc_fence_after();
                        const uint64_t vdesc = tc_smem_desc_sw128(base + OFF_V + sl * TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < BKV / 16; ++k)
                            tc_mma_ts(tmem_base + (uint32_t)(j * 128 + 64), tmem_base + (uint32_t)(j * 128 + 8 * k),
                                      vdesc + (uint64_t)(k * (2048 >> 4)), idesc_o, k != 0 ? 1u : 0u);
                    }
                    tc_commit(bars.o());
                    ATL(it, 2);
                    // once these MMAs have completed, key tile qt-1 of the item (and, after the last query tile, the rest)
                    // is dead: refill its ring slot with the key tile four ahead, and fetch the Q after next
                    mbar_wait(bars.o(), (uint32_t)(it & 1));
                    ATL(it, 3);
                    kl_released = first_kl + (qt + 1 < nqt ? qt + 1 : nqt + 2);
                    ++it;
                    issue_loads();
                    --it;
                    ATL(it, 4);
                }
            }
        }
        __syncwarp();
    } else {
        // ===== softmax warps: a query row is shared by the two warps of its TMEM lane quarter (warp % 4): warp `hf` = warp / 4
        // takes columns [64 hf, 64 hf + 64) of every key tile.  One warp per SM sub-partition left the kernel latency
        // bound (IPC 0.14 in ncu); two interleave their tcgen05.ld / MUFU / tcgen05.st chains. =====
        const int quarter = warp & 3, hf = warp >> 2;
        const int r = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        float* xch = reinterpret_cast<float*>(gbase + OFF_XCH);      // [slot 0..3][hf][row]
        auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;\n" ::"r"(1 + quarter) : "memory"); };
        int it = 0;
        for (int item = blockIdx.x; item < items; item += (int)gridDim.x) {
            const int n = item / NH, h = item % NH;
            for (int qt = 0; qt < nqt; ++qt, ++it) {
                const int q0 = qt * BQ, q = q0 + r, k0 = q0 - BKV;
                // visible band of this row in band columns c = key - k0 (0 .. 383)
                const int lo = max(BKV + r - wl, -k0), hi = min(BKV + r + wr, T - 1 - k0);
                const uint32_t ph = (uint32_t)(it & 1);
                float m0 = -1e30f, m1 = -1e30f, m2 = -1e30f, l0 = 0.f, l1 = 0.f, l2 = 0.f;   // m: row maxima; l: THIS warp's partial sums
#pragma unroll 1
                for (int j = 0; j < NKT; ++j) {
                    // visible columns of this warp's half of tile j, relative to the half: [a, b] within 0..63
                    const int a = max(lo - j * BKV - hf * 64, 0), b = min(hi - j * BKV - hf * 64, 63);
                    mbar_wait(bars.s(j), ph);
                    tc_fence_after();
                    if (warp == 0 && lane == 0) { if (j == 0) ATL(it, 5); else if (j == 1) ATL(it, 9); }
                    uint32_t s[64];
                    int kind[2];     // per 32-column piece: 0 = nobody in the warp sees it (not read), 2 = everybody sees all of it, 1 = mixed
#pragma unroll
                    for (int pc = 0; pc < 2; ++pc) {
                        const bool any = __any_sync(0xffffffffu, a <= pc * 32 + 31 && b >= pc * 32);
                        const bool all = __all_sync(0xffffffffu, a <= pc * 32 && b >= pc * 32 + 31);
                        kind[pc] = all ? 2 : (any ? 1 : 0);
                        if (any) tc_ld_32x32b_x32(lane_addr + (uint32_t)(j * 128 + hf * 64 + pc * 32), *reinterpret_cast<uint32_t(*)[32]>(&s[pc * 32]));
                    }
                    tc_wait_ld();
                    if (warp == 0 && lane == 0 && j == 0) ATL(it, 6);
                    constexpr uint32_t NEG_INF = 0xff800000u;     // masked scores -> -inf (2^-inf = 0 below)
#pragma unroll
                    for (int pc = 0; pc < 2; ++pc) {
                        if (kind[pc] == 1) {
#pragma unroll
                            for (int c = 0; c < 32; ++c) {
                                const int col = pc * 32 + c;
                                s[col] = (col >= a && col <= b) ? s[col] : NEG_INF;
                            }
                        } else if (kind[pc] == 0) {
#pragma unroll
                            for (int c = 0; c < 32; ++c) s[pc * 32 + c] = NEG_INF;
                        }
                    }
                    float mx0 = -1e30f, mx1 = -1e30f, mx2 = -1e30f, mx3 = -1e30f;     // (finite floor: a fully masked row stays well defined)
#pragma unroll
                    for (int c = 0; c < 64; c += 4) {
                        mx0 = fmaxf(mx0, __uint_as_float(s[c]));
                        mx1 = fmaxf(mx1, __uint_as_float(s[c + 1]));
                        mx2 = fmaxf(mx2, __uint_as_float(s[c + 2]));
                        mx3 = fmaxf(mx3, __uint_as_float(s[c + 3]));
                    }
                    float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                    // row maximum over both halves: exchange through shared memory with the partner warp
                    xch[((j & 1) * 2 + hf) * BQ + r] = mx;
                    pair_sync();
                    mx = fmaxf(mx, xch[((j & 1) * 2 + (hf ^ 1)) * BQ + r]);
                    if (warp == 0 && lane == 0 && j == 0) ATL(it, 7);
                    const float mb = mx * scale_log2e;
                    float sum0 = 0.f, sum1 = 0.f;
                    uint32_t pk[32];
#pragma unroll
                    for (int pc = 0; pc < 2; ++pc) {
                        if (kind[pc] != 0) {          // (warp-uniform) a piece nobody sees costs no exponentials: P = 0
#pragma unroll
                            for (int c2 = 0; c2 < 16; ++c2) {
                                // two exponentials per MUFU operation: the arguments (<= 0) are rounded to fp16 -- P is stored
                                // as fp16 anyway -- and ex2.approx.f16x2 returns the packed pair the A operand wants; the row
                                // sum is taken over the ROUNDED probabilities, i.e. over exactly what the tensor core multiplies
                                const __half2 x2 = __floats2half2_rn(fmaf(__uint_as_float(s[pc * 32 + 2 * c2]), scale_log2e, -mb),
                                                                     fmaf(__uint_as_float(s[pc * 32 + 2 * c2 + 1]), scale_log2e, -mb));
                                uint32_t pbits;
                                asm("ex2.approx.f16x2 %0, %1;" : "=r"(pbits) : "r"(*reinterpret_cast<const uint32_t*>(&x2)));
                                const float2 pf = __half22float2(*reinterpret_cast<const __half2*>(&pbits));
                                sum0 += pf.x;
                                sum1 += pf.y;
                                pk[pc * 16 + c2] = pbits;
                            }
                        } else {
#pragma unroll
                            for (int c2 = 0; c2 < 16; ++c2) pk[pc * 16 + c2] = 0u;
                        }
                    }
                    tc_st_32x32b_x32(lane_addr + (uint32_t)(j * 128 + hf * 32), pk);   // P columns of keys 64 hf .. 64 hf + 63
                    const float sum = sum0 + sum1;
                    if (j == 0) { m0 = mx; l0 = sum; } else if (j == 1) { m1 = mx; l1 = sum; } else { m2 = mx; l2 = sum; }
                    tc_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bars.p(j));
                    if (warp == 0 && lane == 0) { if (j == 0) ATL(it, 8); else if (j == 1) ATL(it, 10); else ATL(it, 11); }
                }
                // ===== epilogue: combine the three partial softmaxes of the row; this warp writes head dims [32 hf, 32 hf + 32) =====
                const float M = fmaxf(fmaxf(m0, m1), m2);
                float f0 = m0 <= -1e30f ? 0.f : ex2_approx(fmaxf((m0 - M) * scale_log2e, -126.f));
                float f1 = m1 <= -1e30f ? 0.f : ex2_approx(fmaxf((m1 - M) * scale_log2e, -126.f));
                float f2 = m2 <= -1e30f ? 0.f : ex2_approx(fmaxf((m2 - M) * scale_log2e, -126.f));
                const float Lpart = fmaf(l0, f0, fmaf(l1, f1, l2 * f2));
                xch[(2 * 2 + hf) * BQ + r] = Lpart;         // slots 2: the partner's share of the denominator
                pair_sync();
                const float L = Lpart + xch[(2 * 2 + (hf ^ 1)) * BQ + r];
                const float inv = L > 0.f ? 1.0f / L : 0.f;
                f0 *= inv; f1 *= inv; f2 *= inv;
                mbar_wait(bars.o(), ph);
                tc_fence_after();
                if (warp == 0 && lane == 0) ATL(it, 12);
                __half* dst = out + ((size_t)n * T + q) * (size_t)(NH * HD) + h * HD + hf * 32;
                {
                    uint32_t o0[32], o1[32], o2[32];
                    tc_ld_32x32b_x32(lane_addr + (uint32_t)(0 * 128 + 64 + hf * 32), o0);
                    tc_ld_32x32b_x32(lane_addr + (uint32_t)(1 * 128 + 64 + hf * 32), o1);
                    tc_ld_32x32b_x32(lane_addr + (uint32_t)(2 * 128 + 64 + hf * 32), o2);
                    tc_wait_ld();
                    // all of this tile's TMEM has been read by this warp: the next tile's MMAs may overwrite it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bars.tfree());
                    if (warp == 0 && lane == 0) ATL(it, 13);
                    if (q < T) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            __half2 hh[4];
#pragma unroll
                            for (int p = 0; p < 4; ++p) {
                                const int c = g * 8 + 2 * p;
                                const float v0 = __uint_as_float(o0[c]) * f0 + __uint_as_float(o1[c]) * f1 + __uint_as_float(o2[c]) * f2;
                                const float v1 = __uint_as_float(o0[c + 1]) * f0 + __uint_as_float(o1[c + 1]) * f1 + __uint_as_float(o2[c + 1]) * f2;
                                hh[p] = __floats2half2_rn(v0, v1);
                            }
                            *reinterpret_cast<uint4*>(dst + g * 8) = *reinterpret_cast<const uint4*>(hh);
                        }
                    }
                    if (warp == 0 && lane == 0) ATL(it, 14);
                }
            }
        }
    }
#undef ATL

    tc_fence_before();
    __syncthreads();
    if (warp == SM_WARPS) tc_dealloc(tmem_base, TMEM_COLS);
}


// ---- second version: sixteen softmax warps, one output accumulator ------------------------------------------------------
// The kernel above spends ~10 k cycles per query tile against ~1.5 k of tensor work: each of its eight softmax warps walks
// through load -> max -> exchange -> exp -> store three times per tile with nothing to overlap, the three partial outputs are
// read back and combined in registers, and the next tile's Q K^T cannot start before that has happened (S_j and O_j share
// columns).  Here
//   * the row maximum is taken over the WHOLE band first (pass 1 reads S_0..S_2, one exchange per tile), so all three
//     P_j V_j products accumulate into ONE 64-column O: a third of the read-back, no rescaling, and O has its own columns
//     [384, 448): the next tile's Q K^T is issued as soon as this tile's MMAs have completed, while the softmax warps are
//     still writing this tile's output;
//   * the row sums come from the tensor core as well: P_j times a 128 x 16 block of ones accumulates into columns
//     [448, 464) -- exactly the sum of the fp16 probabilities the P V product sees, no unpack / add per element, no second
//     exchange;
//   * sixteen softmax warps, four per TMEM lane quarter, 32 of a key tile's 128 columns each (four warps per scheduler
//     instead of two); warp c keeps its P in the first 16 of ITS OWN 32 columns of S_j, so no warp overwrites scores another
//     one has yet to read in pass 2.
namespace v2 {

constexpr int SM_WARPS = 16, CSPLIT = 4;
constexpr int THREADS = (SM_WARPS + 1) * 32;
constexpr uint32_t OFF_XCH = OFF_BARS + 256;                             // float [2][CSPLIT][BQ] row maxima
constexpr uint32_t OFF_ONES = (OFF_XCH + 2 * CSPLIT * BQ * 4 + 1023u) & ~1023u;   // 16 rows x 128 B of fp16 ones (B operand of the row sums)
constexpr uint32_t SMEM_BYTES = OFF_ONES + 2048 + 1024;
constexpr uint32_t COL_O = 384, COL_L = 448;

__global__ void __launch_bounds__(THREADS, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap map_qkv, __half* __restrict__ out, int N, int T, int NH, int wl,
                     int wr, float scale_log2e, int debug) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    AttnBars bars;
    bars.base = base + OFF_BARS;
    const uint32_t tmem_slot = bars.base + 112;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nqt = (T + BQ - 1) / BQ;
    const int items = N * NH;
    const bool tl = debug != 0 && blockIdx.x == 0;
#define ATL(it_, k_) do { if (tl && (it_) < ATL_TILES) g_attn_timeline[(it_)][(k_)] = clock64(); } while (0)

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_qkv));
        for (int b = 0; b < 2; ++b) mbar_init(bars.qfull(b), 1);
        for (int sl = 0; sl < RING; ++sl) mbar_init(bars.kvfull(sl), 1);
        for (int j = 0; j < NKT; ++j) mbar_init(bars.s(j), 1);
        mbar_init(bars.p(0), SM_WARPS);            // one arrive per softmax warp, after all three P_j
        mbar_init(bars.o(), 1);
        mbar_init(bars.tfree(), SM_WARPS);
        mbar_fence_init();
    }
    for (int i = tid; i < 512; i += THREADS) reinterpret_cast<uint32_t*>(gbase + OFF_ONES)[i] = 0x3C003C00u;   // fp16 1.0 pairs
    fence_proxy_async_smem();
    if (warp == SM_WARPS) tc_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + OFF_BARS + 112);

    if (warp == SM_WARPS) {
        // ===== TMA producer + MMA issuer (one thread) =====
        if (elect_one_sync()) {
            constexpr uint32_t idesc_s = tc_idesc_f16(BQ, BKV);
            constexpr uint32_t idesc_o = tc_idesc_f16(BQ, HD) | (1u << 16);   // bit 16: B (V) is MN-major
            constexpr uint32_t idesc_l = tc_idesc_f16(BQ, 16);               // B = ones, K-major
            const uint64_t ones_desc = tc_smem_desc_sw128(base + OFF_ONES);
            int it = 0, kl_issued = 0, q_issued = 0, kl_released = 0;
            const int my_items = blockIdx.x < items ? (items - blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
            const int total_q = my_items * nqt, total_k = my_items * (nqt + 2);

            auto issue_loads = [&]() {
                while (kl_issued < total_k && kl_issued < kl_released + RING) {
                    const int li = kl_issued / (nqt + 2), kt = kl_issued % (nqt + 2) - 1;
                    const int item = blockIdx.x + li * (int)gridDim.x, n = item / NH, h = item % NH;
                    const int sl = kl_issued & (RING - 1);
                    mbar_expect_tx(bars.kvfull(sl), 2 * TILE_BYTES);
                    tma_load_3d(base + OFF_K + sl * TILE_BYTES, &map_qkv, bars.kvfull(sl), NH * HD + h * HD, kt * BKV, n);
                    tma_load_3d(base + OFF_V + sl * TILE_BYTES, &map_qkv, bars.kvfull(sl), 2 * NH * HD + h * HD, kt * BKV, n);
                    ++kl_issued;
                }
                while (q_issued < total_q && q_issued < it + 2) {
                    const int li = q_issued / nqt, qt = q_issued % nqt;
                    const int item = blockIdx.x + li * (int)gridDim.x, n = item / NH, h = item % NH;
                    mbar_expect_tx(bars.qfull(q_issued & 1), TILE_BYTES);
                    tma_load_3d(base + OFF_Q + (q_issued & 1) * TILE_BYTES, &map_qkv, bars.qfull(q_issued & 1), h * HD, qt * BQ, n);
                    ++q_issued;
                }
            };

            issue_loads();
            for (int li = 0; li < my_items; ++li) {
                const int first_kl = li * (nqt + 2);
                for (int qt = 0; qt < nqt; ++qt, ++it) {
                    // S_j = Q K_j^T: the previous tile's MMAs have completed (wait on o() below), so S / P are free
                    mbar_wait(bars.qfull(it & 1), (uint32_t)((it >> 1) & 1));
                    tc_fence_after();
                    ATL(it, 0);
                    const uint64_t qdesc = tc_smem_desc_sw128(base + OFF_Q + (it & 1) * TILE_BYTES);
#pragma unroll
                    for (int j = 0; j < NKT; ++j) {
                        const int kl = first_kl + qt + j, sl = kl & (RING - 1);
                        mbar_wait(bars.kvfull(sl), (uint32_t)((kl / RING) & 1));
                        tc_fence_after();
                        const uint64_t kdesc = tc_smem_desc_sw128(base + OFF_K + sl * TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k)
                            tc_mma_ss(tmem_base + (uint32_t)(j * 128), qdesc + 2u * k, kdesc + 2u * k, idesc_s, k != 0 ? 1u : 0u);
                        tc_commit(bars.s(j));
                    }
                    ATL(it, 1);
                    // the loads the previous tile's completion made room for: off the critical path (the softmax warps are busy
                    // with S for the next few thousand cycles; the ring holds this tile's keys and the next one's already)
                    issue_loads();
                    ATL(it, 4);
                    // O += P_j V_j and L += P_j 1: P of keys [32 c, 32 c + 32) sits in columns [32 c, 32 c + 16) of S_j
                    mbar_wait(bars.p(0), (uint32_t)(it & 1));
                    if (it > 0) mbar_wait(bars.tfree(), (uint32_t)((it - 1) & 1));    // the previous tile's O / L have been read
                    tc_fence_after();
#pragma unroll
                    for (int j = 0; j < NKT; ++j) {
                        const int kl = first_kl + qt + j, sl = kl & (RING - 1);
                        const uint64_t vdesc = tc_smem_desc_sw128(base + OFF_V + sl * TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < BKV / 16; ++k) {
                            const uint32_t a = tmem_base + (uint32_t)(j * 128 + (k >> 1) * 32 + (k & 1) * 8);
                            tc_mma_ts(tmem_base + COL_O, a, vdesc + (uint64_t)(k * (2048 >> 4)), idesc_o, (j | k) != 0 ? 1u : 0u);
                            tc_mma_ts(tmem_base + COL_L, a, ones_desc, idesc_l, (j | k) != 0 ? 1u : 0u);
                        }
                    }
                    tc_commit(bars.o());
                    ATL(it, 2);
                    mbar_wait(bars.o(), (uint32_t)(it & 1));
                    ATL(it, 3);
                    kl_released = first_kl + (qt + 1 < nqt ? qt + 1 : nqt + 2);
                    if (qt + 1 == nqt) issue_loads();      // the next item's first key tiles: nothing else can load them in time
                }
            }
        }
        __syncwarp();
    } else {
        // ===== softmax warps: row r = 32 quarter + lane; warp c = warp / 4 takes columns [32 c, 32 c + 32) of every key tile =====
        const int quarter = warp & 3, cq = warp >> 2;
        const int r = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        float* xch = reinterpret_cast<float*>(gbase + OFF_XCH);      // [parity][c][row]
        constexpr uint32_t NEG_INF = 0xff800000u;
        int it = 0;
        for (int item = blockIdx.x; item < items; item += (int)gridDim.x) {
            const int n = item / NH, h = item % NH;
            for (int qt = 0; qt < nqt; ++qt, ++it) {
                const int q0 = qt * BQ, q = q0 + r, k0 = q0 - BKV;
                const int lo = max(BKV + r - wl, -k0), hi = min(BKV + r + wr, T - 1 - k0);   // visible band columns (0 .. 383)
                const uint32_t ph = (uint32_t)(it & 1);
                // ---- pass 1: row maximum over the visible part of this warp's 3 x 32 columns ----
                float mx = -1e30f;       // (finite floor: a fully masked row stays well defined)
                int kinds = 0;           // 2 bits per key tile: 0 = no lane of the warp sees the piece, 2 = all see all of it, 1 = mixed
#pragma unroll 1
                for (int j = 0; j < NKT; ++j) {
                    const int a = max(lo - j * BKV - cq * 32, 0), b = min(hi - j * BKV - cq * 32, 31);
                    const bool any = __any_sync(0xffffffffu, a <= b);
                    const bool all = __all_sync(0xffffffffu, a == 0 && b == 31);
                    const int kind = all ? 2 : (any ? 1 : 0);
                    kinds |= kind << (2 * j);
                    mbar_wait(bars.s(j), ph);
                    if (warp == 0 && lane == 0 && j == 0) ATL(it, 5);
                    if (kind == 0) continue;
                    tc_fence_after();
                    uint32_t s[32];
                    tc_ld_32x32b_x32(lane_addr + (uint32_t)(j * 128 + cq * 32), s);
                    tc_wait_ld();
                    float m0 = -1e30f, m1 = -1e30f, m2 = -1e30f, m3 = -1e30f;
                    if (kind == 2) {
#pragma unroll
                        for (int c = 0; c < 32; c += 4) {
                            m0 = fmaxf(m0, __uint_as_float(s[c]));
                            m1 = fmaxf(m1, __uint_as_float(s[c + 1]));
                            m2 = fmaxf(m2, __uint_as_float(s[c + 2]));
                            m3 = fmaxf(m3, __uint_as_float(s[c + 3]));
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; c += 4) {
                            m0 = fmaxf(m0, (c >= a && c <= b) ? __uint_as_float(s[c]) : -1e30f);
                            m1 = fmaxf(m1, (c + 1 >= a && c + 1 <= b) ? __uint_as_float(s[c + 1]) : -1e30f);
                            m2 = fmaxf(m2, (c + 2 >= a && c + 2 <= b) ? __uint_as_float(s[c + 2]) : -1e30f);
                            m3 = fmaxf(m3, (c + 3 >= a && c + 3 <= b) ? __uint_as_float(s[c + 3]) : -1e30f);
                        }
                    }
                    mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
                }
                if (warp == 0 && lane == 0) ATL(it, 6);
                xch[((int)ph * CSPLIT + cq) * BQ + r] = mx;
                asm volatile("bar.sync %0, 128;\n" ::"r"(1 + quarter) : "memory");    // the four warps of this lane quarter
#pragma unroll
                for (int c = 0; c < CSPLIT; ++c) mx = fmaxf(mx, xch[((int)ph * CSPLIT + c) * BQ + r]);
                if (warp == 0 && lane == 0) ATL(it, 7);
                const float mb = mx * scale_log2e;
                // ---- pass 2: P = 2^(s * scale - m), fp16 pairs, into the first 16 of this warp's own columns ----
#pragma unroll 1
                for (int j = 0; j < NKT; ++j) {
                    const int kind = (kinds >> (2 * j)) & 3;
                    uint32_t pk[16];
                    if (kind != 0) {
                        const int a = max(lo - j * BKV - cq * 32, 0), b = min(hi - j * BKV - cq * 32, 31);
                        uint32_t s[32];
                        tc_ld_32x32b_x32(lane_addr + (uint32_t)(j * 128 + cq * 32), s);
                        tc_wait_ld();
                        if (kind == 1) {
#pragma unroll
                            for (int c = 0; c < 32; ++c) s[c] = (c >= a && c <= b) ? s[c] : NEG_INF;    // 2^-inf = 0
                        }
#pragma unroll
                        for (int c2 = 0; c2 < 16; ++c2) {
                            // arguments (<= 0) rounded to fp16 -- P is stored as fp16 anyway -- two exponentials per MUFU operation
                            const __half2 x2 = __floats2half2_rn(fmaf(__uint_as_float(s[2 * c2]), scale_log2e, -mb),
                                                                 fmaf(__uint_as_float(s[2 * c2 + 1]), scale_log2e, -mb));
                            asm("ex2.approx.f16x2 %0, %1;" : "=r"(pk[c2]) : "r"(*reinterpret_cast<const uint32_t*>(&x2)));
                        }
                    } else {
#pragma unroll
                        for (int c2 = 0; c2 < 16; ++c2) pk[c2] = 0u;
                    }
                    tc_st_32x32b_x16(lane_addr + (uint32_t)(j * 128 + cq * 32), pk);
                }
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars.p(0));
                if (warp == 0 && lane == 0) ATL(it, 8);
                // ---- epilogue: out[q][16 c .. 16 c + 16) of this head = O / L ----
                mbar_wait(bars.o(), ph);
                tc_fence_after();
                if (warp == 0 && lane == 0) ATL(it, 12);
                uint32_t o[16], lsum;
                tc_ld_32x32b_x16(lane_addr + COL_O + (uint32_t)(cq * 16), o);
                tc_ld_32x32b_x1(lane_addr + COL_L, lsum);
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars.tfree());
                if (warp == 0 && lane == 0) ATL(it, 13);
                if (q < T) {
                    const float L = __uint_as_float(lsum);
                    const float inv = L > 0.f ? 1.0f / L : 0.f;
                    __half* dst = out + ((size_t)n * T + q) * (size_t)(NH * HD) + h * HD + cq * 16;
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        __half2 hh[4];
#pragma unroll
                        for (int p = 0; p < 4; ++p)
                            hh[p] = __floats2half2_rn(__uint_as_float(o[g * 8 + 2 * p]) * inv, __uint_as_float(o[g * 8 + 2 * p + 1]) * inv);
                        *reinterpret_cast<uint4*>(dst + g * 8) = *reinterpret_cast<const uint4*>(hh);
                    }
                }
                if (warp == 0 && lane == 0) ATL(it, 14);
            }
        }
    }
#undef ATL

    tc_fence_before();
    __syncthreads();
    if (warp == SM_WARPS) tc_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace v2

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace

// the tcgen05 kernel covers windows whose band fits three key tiles; anything else stays on the mma.sync kernel
bool attention_tc_supported(int head_dim, int wl, int wr) { return head_dim == HD && wl >= 0 && wr >= 0 && wl <= BKV && wr <= BKV; }

// qkv [N][T][3][NH][64] (rotary already applied to q, k) -> out [N][T][NH*64]
int launch_attention_tc(const __half* qkv, __half* out, int N, int T, int NH, int wl, int wr, cudaStream_t stream) {
    B200_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0, "attention: operands must be 16-byte aligned");
    EncodeTiledFn fn = encode_fn();
    B200_REQUIRE(fn != nullptr, "attention: cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap map;
    const cuuint64_t width = (cuuint64_t)3 * NH * HD;
    cuuint64_t dims[3] = {width, (cuuint64_t)T, (cuuint64_t)N};
    cuuint64_t strides[2] = {width * 2, width * 2 * (cuuint64_t)T};
    cuuint32_t box[3] = {(cuuint32_t)HD, (cuuint32_t)BQ, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(qkv), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, "attention: cuTensorMapEncodeTiled failed (%d)", (int)r);
    B200_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    int dev = 0, sms = 0;
    B200_CHECK_CUDA(cudaGetDevice(&dev));
    B200_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int items = N * NH;
    const int grid = items < sms ? items : sms;
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)HD);
    const char* dbg = getenv("B200_ATTN_DEBUG");
    static int version = 0;     // B200_ATTN_TC=1: the first version (eight softmax warps, per-key-tile maxima)
    if (version == 0) {
        const char* e = getenv("B200_ATTN_TC");
        version = (e && e[0] == '1') ? 1 : 2;
    }
    if (version == 1) {
        attention_tc_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(map, out, N, T, NH, wl, wr, scale_log2e, dbg ? atoi(dbg) : 0);
    } else {
        B200_CHECK_CUDA(cudaFuncSetAttribute(v2::attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v2::SMEM_BYTES));
        v2::attention_tc2_kernel<<<grid, v2::THREADS, v2::SMEM_BYTES, stream>>>(map, out, N, T, NH, wl, wr, scale_log2e,
                                                                                 dbg ? atoi(dbg) : 0);
    }
    B200_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int copy_attention_timeline(long long* host_out, int max_tiles) {
    const int n = max_tiles < ATL_TILES ? max_tiles : ATL_TILES;
    B200_CHECK_CUDA(cudaDeviceSynchronize());
    B200_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_attn_timeline, sizeof(long long) * 16 * n));
    return n;
}
