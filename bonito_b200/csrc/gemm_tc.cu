// C[M,N] = epilogue(A[M,K] * B[N,K]^T) on the 5th-generation tensor cores (sm_100a).
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0     TMA producer: cp.async.bulk.tensor 2D tiles (SWIZZLE_128B) of A (128 x 64) and B (BN x 64)
//              into a 4-stage shared-memory ring, completion on mbarriers
//   warp 1     MMA issuer: one lane issues tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN, K=16),
//              accumulating in TMEM; tcgen05.commit releases ring slots / publishes accumulators
//   warps 2-13 epilogue, three warps per 32-lane TMEM quarter taking every third 32-column block: tcgen05.ld (32 lanes x
//              32 columns) -> bias -> fp16 rounding -> activation (or the fused SwiGLU product) -> row-remapped
//              16-byte stores.  TMEM holds two accumulator stages (2 x BN columns) so the epilogue of tile i overlaps
//              the mainloop of tile i+1 (K is only 304..512 here, so the epilogue is as long as the mainloop).
// Used for: the strided conv of the encoder (overlapping-row view of the conv-stem output, lda < K),
// the LSTM input projections and the LinearCRFEncoder (+Clamp) -- reference call sites
// bonito/nn.py:226,235-241 (Conv1d), :366-370 (LSTM W_ih), :283-298 + :59-67 (Linear, Clamp).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "tc_common.cuh"

namespace {

constexpr int BM = 128, BK = 64, STAGES = 4;
// Three epilogue warps per TMEM lane quarter: with two, the MMA warp spent most of its waiting on `tempty` (ncu source page:
// the epilogue's dependent tcgen05.ld -> convert -> st.shared -> shfl / ld.shared -> st.global chain is latency bound, issue
// slots 34 % busy), tensor pipe 48 % active.
constexpr int EPI_SETS = 3;
constexpr int EPI_WARPS = 4 * EPI_SETS;
constexpr int THREADS = (2 + EPI_WARPS) * 32;   // TMA warp, MMA warp, epilogue warps

template <int BN>
struct TcSmem {
    static constexpr uint32_t kA = BM * BK * 2;         // 16 KB
    static constexpr uint32_t kB = BN * BK * 2;         // 16 / 32 KB
    static constexpr uint32_t kStage = kA + kB;
    static constexpr uint32_t kEpi = STAGES * kStage;             // 8 epilogue warps x (32 rows x 64 B) transpose buffers
    static constexpr uint32_t kBias = kEpi + EPI_WARPS * 2048;    // 512 B per warp: the tile's bias slice
    static constexpr uint32_t kBars = kBias + EPI_WARPS * 512;    // mbarriers after the buffers
    static constexpr uint32_t kTileQ = kBars + 160;               // dynamic tile queue: TQ mbarriers + TQ ints
    static constexpr uint32_t kTotal = kBars + 384 + 1024;   // + alignment slack
};

// ---- PTX wrappers not shared with the other tcgen05 kernels -----------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
            "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    tc_ld_32x32b_x32(taddr, v);
    tc_wait_ld();
}

// ---- dynamic tile scheduling -------------------------------------------------------------------------------------
// The kernels are persistent (one CTA per SM), but the engine runs them next to recurrent clusters that hold 66-132 of the
// 148 SMs for milliseconds: with a static tile -> CTA assignment the CTAs that cannot be placed before those clusters
// retire would hold back the whole GEMM.  Tiles are therefore handed out by a ticket counter in global memory
// (atomicAdd by the TMA-producer thread, one tile ahead); the ticket travels to the MMA and epilogue warps through a
// small shared-memory queue.  CTAs that start late find the counter exhausted and leave at once.  The counters come
// from a zero-initialised pool; the last CTA to finish resets the ones its launch used.
constexpr int TQ = 8;   // queue depth; the producer is never more than 3 tiles ahead of the slowest epilogue warp

struct TileQueue {
    uint32_t bars;          // TQ mbarriers (shared-memory address)
    volatile int* tiles;    // TQ tile indices (generic pointer)
    __device__ __forceinline__ void publish(int q, int tile) const {   // producer thread
        tiles[q & (TQ - 1)] = tile;
        mbar_arrive(bars + 8u * (uint32_t)(q & (TQ - 1)));             // release: the store above is visible to the waiters
    }
    __device__ __forceinline__ int take(int q) const {                 // any consumer thread
        mbar_wait(bars + 8u * (uint32_t)(q & (TQ - 1)), (uint32_t)((q / TQ) & 1));
        return tiles[q & (TQ - 1)];
    }
};

// last CTA out resets the `n_ctr` ticket counters and the exit counter behind them
__device__ __forceinline__ void release_tile_counters(int* ctr, int n_ctr) {
    const int done = atomicAdd(&ctr[n_ctr], 1);
    if (done == (int)gridDim.x - 1) {
        for (int i = 0; i <= n_ctr; ++i) ctr[i] = 0;
        __threadfence();
    }
}

// One warp copies the BN bias values of column block `nb` into its shared-memory slice (zeros when there is no bias).
template <int BN>
__device__ __forceinline__ void stage_bias(__half* sbias, const __half* __restrict__ bias, int nb, int N, int lane) {
    if (lane < BN / 8) {
        const int gn = nb * BN + lane * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (bias != nullptr && gn < N) v = __ldg(reinterpret_cast<const uint4*>(bias + gn));
        reinterpret_cast<uint4*>(sbias)[lane] = v;
    }
    __syncwarp();
}

// Epilogue of one 128 x BN accumulator tile.  EPI_SETS warps share each 32-lane TMEM quarter (= 32 output rows): warp set
// `set` takes every other 32-column block, so the two sets drain one accumulator in half the time (the kernels were
// epilogue-bound with a single set: tile period 5.6k cycles against a 3.4k-cycle mainloop).  A lane owns one row; a block
// of 32 columns is converted, transposed through a swizzled 32 x 64 B shared-memory buffer and written so that 4 lanes
// cover 64 contiguous bytes of one output row (two full sectors instead of 32 scattered half-sectors per instruction).
template <int V>
struct ActC { static constexpr int value = V; };

template <int BN, int I8 = 0>
__device__ __forceinline__ void epilogue_tile(uint32_t tmem_acc, unsigned char* tbuf, const __half* sbias,
                                              __half* __restrict__ C, long long ldc, int M, int N, int mb, int nb,
                                              int quarter, int set, int lane, const GemmEpilogue& ep,
                                              const float* sscale = nullptr) {
    const int gm = mb * BM + quarter * 32 + lane;
    const long long orow = (gm < M) ? map_row(ep.map, gm) : -1;
    const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
    const int cb_shift = (ep.cb_width > 0 && (ep.cb_width & (ep.cb_width - 1)) == 0) ? __ffs(ep.cb_width) - 1 : -1;
    // 64-B rows: the row's parity picks the half of a 128-B line, (row >> 1) & 3 permutes the four 16-B slots
    auto stage = [&](int g, const __half2 (&packed)[4]) {
        *reinterpret_cast<uint4*>(tbuf + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(packed);
    };
    auto flush = [&](int gcol, int ncols) {   // staged 32 columns -> C[:, gcol .. gcol+31]; ncols = valid output width
        __syncwarp();
        const int chunk = lane & 3;
        const bool col_ok = gcol + chunk * 8 < ncols;   // widths are multiples of 8
        // column-block remap (cb_width is a multiple of 32, so a 32-column block never straddles two column blocks)
        long long row_add = 0;
        int ocol = gcol;
        if (ep.cb_width > 0) {
            const int cb = cb_shift >= 0 ? (gcol >> cb_shift) : gcol / ep.cb_width;
            row_add = (long long)cb * ep.cb_rows;
            ocol = gcol - cb * ep.cb_width;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = 8 * i + (lane >> 2);
            const long long r = __shfl_sync(0xffffffffu, orow, row);
            const uint4 val = *reinterpret_cast<const uint4*>(tbuf + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
            if (r >= 0 && col_ok) *reinterpret_cast<uint4*>(C + (r + row_add) * ldc + ocol + chunk * 8) = val;
        }
        __syncwarp();
    };
    if (ep.act == B200_ACT_SWIGLU) {
        // columns [c0, c0+32) = y, [c0+32, c0+64) = gate of the same 32 features -> 32 outputs per row
#pragma unroll 1
        for (int c0 = set * 64; c0 < BN; c0 += 64 * EPI_SETS) {
            const int gn0 = nb * BN + c0;
            if (gn0 >= N) break;  // warp-uniform
            uint32_t vy[32], vg[32];
            tc_ld_32x32b_x32(taddr + c0, vy);
            tc_ld_32x32b_x32(taddr + c0 + 32, vg);
            tc_wait_ld();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                __half2 packed[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float y0 = round_f16(__uint_as_float(vy[g * 8 + 2 * p])), y1 = round_f16(__uint_as_float(vy[g * 8 + 2 * p + 1]));
                    const float g0 = round_f16(__uint_as_float(vg[g * 8 + 2 * p])), g1 = round_f16(__uint_as_float(vg[g * 8 + 2 * p + 1]));
                    packed[p] = __floats2half2_rn(g0 * y0 * rcp_approx(1.0f + __expf(-g0)), g1 * y1 * rcp_approx(1.0f + __expf(-g1)));
                }
                stage(g, packed);
            }
            flush(gn0 >> 1, N >> 1);
        }
        return;
    }
    // The activation code is a kernel argument; dispatching on it per 32-column block (not per element) keeps the
    // conversion loop branch-free: with the switch inside, the CRF (clamp / scale) and convolution (swish / tanh) GEMMs
    // ran their epilogue at a third of the speed of the plain one (340-450 vs 880-1270 TFLOP/s).
    auto block = [&](auto act_c) {
        constexpr int ACT = decltype(act_c)::value;
#pragma unroll 1
        for (int c0 = set * 32; c0 < BN; c0 += 32 * EPI_SETS) {
            const int gn0 = nb * BN + c0;
            if (gn0 >= N) break;  // warp-uniform
            uint32_t v[32];
            tc_ld32(taddr + c0, v);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                __half2 packed[4];
                // bias slice of this tile, staged in shared memory by stage_bias() (a global load here would expose
                // its full latency 32 times per tile: it was 40 % of the epilogue's stall samples)
                const uint4 braw = *reinterpret_cast<const uint4*>(sbias + c0 + g * 8);
                const __half2* bh = reinterpret_cast<const __half2*>(&braw);
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    float x0, x1;
                    if (I8) {   // s32 accumulators of the int8 product, de-quantised with the column's scale
                        x0 = fmaf((float)(int)v[g * 8 + 2 * p], sscale[c0 + g * 8 + 2 * p], __low2float(bh[p]));
                        x1 = fmaf((float)(int)v[g * 8 + 2 * p + 1], sscale[c0 + g * 8 + 2 * p + 1], __high2float(bh[p]));
                    } else {
                        x0 = __uint_as_float(v[g * 8 + 2 * p]) + __low2float(bh[p]);
                        x1 = __uint_as_float(v[g * 8 + 2 * p + 1]) + __high2float(bh[p]);
                    }
                    if (ACT == B200_ACT_NONE)
                        packed[p] = __floats2half2_rn(x0, x1);
                    else
                        packed[p] = __floats2half2_rn(apply_act_f16(x0, ACT, ep.lo, ep.hi), apply_act_f16(x1, ACT, ep.lo, ep.hi));
                }
                stage(g, packed);
            }
            flush(gn0, N);
        }
    };
    switch (ep.act) {
        case B200_ACT_SWISH: block(ActC<B200_ACT_SWISH>()); break;
        case B200_ACT_TANH: block(ActC<B200_ACT_TANH>()); break;
        case B200_ACT_CLAMP: block(ActC<B200_ACT_CLAMP>()); break;
        case B200_ACT_SCALE: block(ActC<B200_ACT_SCALE>()); break;
        case B200_ACT_TANH_SCALE: block(ActC<B200_ACT_TANH_SCALE>()); break;
        default: block(ActC<B200_ACT_NONE>()); break;
    }
}

constexpr uint32_t SMEM_LIMIT = 227 * 1024;
static_assert(TcSmem<256>::kTotal <= SMEM_LIMIT, "streaming kernel: shared memory");

template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               __half* __restrict__ C, long long ldc, int M, int N, int K, GemmEpilogue ep, int* __restrict__ ctr) {
    using S = TcSmem<BN>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-B alignment
    const uint32_t bars = base + S::kBars;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * STAGES + 2 + s); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
    unsigned char* gen_base = smem_raw + (base - smem_u32(smem_raw));
    TileQueue tq;
    tq.bars = base + S::kTileQ;
    tq.tiles = reinterpret_cast<volatile int*>(gen_base + S::kTileQ + 8 * TQ);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blocks = (M + BM - 1) / BM, n_blocks = (N + BN - 1) / BN;
    const int tiles = m_blocks * n_blocks;
    const int k_blocks = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_a));
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_b));
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), EPI_WARPS);  // one arrive per epilogue warp
        }
        for (int s = 0; s < TQ; ++s) mbar_init(tq.bars + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tmem_slot),
                     "r"((uint32_t)(2 * BN)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + S::kBars + 8u * (2 * STAGES + 4));

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one_sync()) {
            int stage = 0;
            uint32_t phase = 0;
            int tile = atomicAdd(ctr, 1);
            for (int q = 0;; ++q) {
                tq.publish(q, tile < tiles ? tile : -1);
                if (tile >= tiles) break;
                const int next = atomicAdd(ctr, 1);        // its latency hides behind the loads below
                const int mb = tile / n_blocks, nb = tile % n_blocks;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), S::kStage);
                    tma_load_2d(base + stage * S::kStage, &map_a, full_bar(stage), kb * BK, mb * BM);
                    tma_load_2d(base + stage * S::kStage + S::kA, &map_b, full_bar(stage), kb * BK, nb * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tile = next;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one_sync()) {
            // instruction descriptor: D=f32, A=B=f16, both K-major, N at [17,23), M at [24,29)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int q = 0; tq.take(q) >= 0; ++q) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint64_t adesc = tc_smem_desc_sw128(base + stage * S::kStage);
                    const uint64_t bdesc = tc_smem_desc_sw128(base + stage * S::kStage + S::kA);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // +32 B per K=16 step inside the 128-B swizzle atom (encoded >> 4)
                        tc_mma_ss(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(empty_bar(stage));  // slot free once these MMAs have read it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull_bar(acc));  // accumulator complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps (TMEM lanes 32*(warp%4) .. +31) =====
        const int quarter = warp & 3, ew = warp - 2, set = ew >> 2;
        unsigned char* tbuf = gen_base + S::kEpi + ew * 2048;
        __half* sbias = reinterpret_cast<__half*>(gen_base + S::kBias + ew * 512);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int q = 0;; ++q) {
            const int tile = tq.take(q);
            if (tile < 0) break;
            const int mb = tile / n_blocks, nb = tile % n_blocks;
            stage_bias<BN>(sbias, ep.bias, nb, N, lane);   // before the wait: its latency hides behind the mainloop
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            epilogue_tile<BN>(tmem_base + (uint32_t)(acc * BN), tbuf, sbias, C, ldc, M, N, mb, nb, quarter, set, lane, ep);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base),
                     "r"((uint32_t)(2 * BN)));
    }
    if (threadIdx.x == 0) release_tile_counters(ctr, 1);
}

// ---- where the cycles of the weight-stationary kernel go (B200_GEMM_DEBUG=1, scripts/gemm_profile.py) -----------------------
// per CTA: [0] lifetime, [1] producer waiting for a free ring slot, [2] MMA thread waiting for A, [3] MMA thread waiting for a
// free accumulator, [4] epilogue warp 2 waiting for an accumulator, [5] epilogue warp 2 inside epilogue_tile, [6] tiles
__device__ long long g_gemm_prof[160 * 8];
static int gemm_debug() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("B200_GEMM_DEBUG");
        mode = e ? atoi(e) : 0;      // 1: counters; +2: A tiles loaded for the first row block only (no L2 traffic for A);
    }                                //             +4: epilogue without global stores; +8: no epilogue work at all (ablations, wrong results)
    return mode;
}

// ---- weight-stationary variant ---------------------------------------------------------------------
// The streaming kernel above re-reads the B tile (BN x K weights, 196 KB at BN=256, K=384) from L2 for every 128 rows of
// A: 294 KB of L2 traffic per tile made the LSTM input projection L2-bandwidth bound (~5.4 TB/s of L2 reads).  Here a
// CTA is bound to ONE column block of B, loads it once into shared memory (<= 144 KB) and streams only A tiles through
// a 3-stage ring: 98 KB of L2 traffic per tile.
constexpr int WS_KB = 6;             // K <= 384 fp16 (K <= 768 int8: a 128-byte row holds 64 halves or 128 bytes)
constexpr int PAIR_DEFAULT = 0;      // cta_group::2 pair kernels: B200_GEMM_PAIR=1 or impl = B200_GEMM_TCGEN05_PAIR (see launch_gemm_tc)

// INT8 variant (I8 = 1; the reference's --quantize path runs koi's int8 LSTM, bonito/crf/model.py:245): A and B are int8 with
// 128-element (128-byte) K blocks, tcgen05.mma kind::i8 accumulates s32 in tensor memory, the epilogue multiplies by a
// per-column scale.  The resident B block shrinks to half, which pays for a 6-stage A ring (3 for fp16): the fp16 kernel is
// starved by the bytes it can keep in flight.
template <int BN, int I8 = 0>
struct WsSmem {
    static constexpr int kStages = I8 ? 6 : 3;
    static constexpr int kKb = I8 ? WS_KB / 2 : WS_KB;
    static constexpr uint32_t kBres = 0;                                  // [kKb][BN rows][128 B] SWIZZLE_128B
    static constexpr uint32_t kRing = kKb * BN * 128;                     // kStages x (128 rows x 128 B)
    static constexpr uint32_t kEpi = kRing + kStages * BM * BK * 2;       // 2 KB transpose buffer per warp
    static constexpr uint32_t kBias = kEpi + EPI_WARPS * 2048;            // 512 B bias slice copy per warp
    static constexpr uint32_t kScale = kBias + EPI_WARPS * 512;           // I8: 8 x 1 KB column-scale copies (float)
    static constexpr uint32_t kBars = kScale + (I8 ? EPI_WARPS * 1024 : 0);
    static constexpr uint32_t kTileQ = kBars + 160;                       // dynamic tile queue: TQ mbarriers + TQ ints
    static constexpr uint32_t kTotal = kBars + 384 + 1024;
};

static_assert(WsSmem<192, 0>::kTotal <= SMEM_LIMIT && WsSmem<192, 1>::kTotal <= SMEM_LIMIT, "weight-stationary kernel: shared memory");

template <int BN, int I8 = 0>
__global__ void __launch_bounds__(THREADS, 1)
gemm_ws_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               __half* __restrict__ C, long long ldc, int M, int N, int K, GemmEpilogue ep, int* __restrict__ ctr,
               const float* __restrict__ col_scale, int prof) {
    using S = WsSmem<BN, I8>;
    const long long t_start = prof ? clock64() : 0;
    long long* pr = g_gemm_prof + (blockIdx.x % 160) * 8;
    constexpr int WS_STAGES = S::kStages;
    constexpr int KE = I8 ? 2 * BK : BK;          // K elements per 128-byte block
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + S::kBars;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (WS_STAGES + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * WS_STAGES + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * WS_STAGES + 2 + s); };
    const uint32_t bres_bar = bars + 8u * (2 * WS_STAGES + 4);
    const uint32_t tmem_slot = bars + 8u * (2 * WS_STAGES + 5);
    unsigned char* gen_base = smem_raw + (base - smem_u32(smem_raw));
    TileQueue tq;
    tq.bars = base + S::kTileQ;
    tq.tiles = reinterpret_cast<volatile int*>(gen_base + S::kTileQ + 8 * TQ);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blocks = (M + BM - 1) / BM, n_blocks = (N + BN - 1) / BN;
    const int k_blocks = (K + KE - 1) / KE;
    const int nb = blockIdx.x % n_blocks;                 // this CTA's column block, for its whole life; row blocks by ticket

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_a));
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_b));
        for (int s = 0; s < WS_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), EPI_WARPS);
        }
        mbar_init(bres_bar, 1);
        for (int s = 0; s < TQ; ++s) mbar_init(tq.bars + 8u * s, 1);
        mbar_fence_init();
    }
    if (warp == 1) tc_alloc(tmem_slot, BN <= 128 ? 256 : 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + S::kBars + 8u * (2 * WS_STAGES + 5));

    if (warp == 0) {
        // ===== TMA producer: B once, then A tiles =====
        if (elect_one_sync()) {
            int mb = atomicAdd(&ctr[nb], 1);
            if (mb < m_blocks) {      // a CTA that starts after the column block is exhausted does not even load B
                mbar_expect_tx(bres_bar, (uint32_t)k_blocks * BN * 128);
                for (int kb = 0; kb < k_blocks; ++kb)
                    tma_load_2d(base + S::kBres + kb * (BN * 128), &map_b, bres_bar, kb * KE, nb * BN);
            }
            int stage = 0;
            uint32_t phase = 0;
            long long w_empty = 0;
            for (int q = 0;; ++q) {
                tq.publish(q, mb < m_blocks ? mb : -1);
                if (mb >= m_blocks) break;
                const int next = atomicAdd(&ctr[nb], 1);   // its latency hides behind the loads below
                for (int kb = 0; kb < k_blocks; ++kb) {
                    const long long t0 = prof ? clock64() : 0;
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    if (prof) w_empty += clock64() - t0;
                    if ((prof & 2) && q > 0) {
                        mbar_arrive(full_bar(stage));      // ablation: the slot keeps the bytes it has
                    } else {
                        mbar_expect_tx(full_bar(stage), BM * BK * 2);
                        tma_load_2d(base + S::kRing + stage * (BM * BK * 2), &map_a, full_bar(stage), kb * KE, mb * BM);
                    }
                    if (++stage == WS_STAGES) { stage = 0; phase ^= 1; }
                }
                mb = next;
            }
            if (prof) pr[1] = w_empty;
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one_sync()) {
            // kind::i8 descriptor: D = s32 (2 at [4,6)), A = B = signed 8-bit (1 at [7,10) and [10,13)), K-major
            const uint32_t idesc = I8 ? ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24))
                                      : tc_idesc_f16(BM, BN);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            long long w_full = 0, w_acc = 0;
            for (int q = 0; tq.take(q) >= 0; ++q) {
                if (q == 0) mbar_wait(bres_bar, 0);
                const long long t0 = prof ? clock64() : 0;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                if (prof) w_acc += clock64() - t0;
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    const long long t1 = prof ? clock64() : 0;
                    mbar_wait(full_bar(stage), phase);
                    if (prof) w_full += clock64() - t1;
                    tc_fence_after();
                    const uint64_t adesc = tc_smem_desc_sw128(base + S::kRing + stage * (BM * BK * 2));
                    const uint64_t bdesc = tc_smem_desc_sw128(base + S::kBres + kb * (BN * 128));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {    // 32 bytes of K per instruction: 16 halves or 32 int8
                        if (I8) tc_mma_ss_i8(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        else tc_mma_ss(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(empty_bar(stage));
                    if (++stage == WS_STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (prof) { pr[2] = w_full; pr[3] = w_acc; }
        }
    } else {
        // ===== epilogue warps =====
        const int quarter = warp & 3, ew = warp - 2, set = ew >> 2;
        unsigned char* tbuf = gen_base + S::kEpi + ew * 2048;
        __half* sbias = reinterpret_cast<__half*>(gen_base + S::kBias + ew * 512);
        long long w_tfull = 0, t_epi = 0, n_tiles = 0;
        stage_bias<BN>(sbias, ep.bias, nb, N, lane);       // the column block never changes
        float* sscale = nullptr;
        if (I8) {
            sscale = reinterpret_cast<float*>(gen_base + S::kScale + ew * 1024);
            for (int i = lane; i < BN; i += 32) sscale[i] = (nb * BN + i < N) ? col_scale[nb * BN + i] : 0.f;
            __syncwarp();
        }
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int q = 0;; ++q) {
            const int mb = tq.take(q);
            if (mb < 0) break;
            const long long t0 = prof ? clock64() : 0;
            mbar_wait(tfull_bar(acc), acc_phase);
            const long long t1 = prof ? clock64() : 0;
            tc_fence_after();
            if (!(prof & 8))
                epilogue_tile<BN, I8>(tmem_base + (uint32_t)(acc * BN), tbuf, sbias, C, ldc, (prof & 4) ? 0 : M, N, mb, nb, quarter, set,
                                      lane, ep, sscale);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (prof) { w_tfull += t1 - t0; t_epi += clock64() - t1; ++n_tiles; }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (prof && ew == 0 && lane == 0) { pr[4] = w_tfull; pr[5] = t_epi; pr[6] = n_tiles; }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tc_dealloc(tmem_base, BN <= 128 ? 256 : 512);
    if (threadIdx.x == 0) release_tile_counters(ctr, n_blocks);
    if (prof && threadIdx.x == 0) pr[0] = clock64() - t_start;
}


__device__ __forceinline__ void st_cluster_s32(uint32_t cluster_addr, int v) {
    asm volatile("st.shared::cluster.s32 [%0], %1;\n" ::"r"(cluster_addr), "r"(v) : "memory");
}

// ---- CTA pairs (cta_group::2): 256 x 256 tiles, the resident weight block split across the two SMs of a pair ---------------
// scripts/gemm_profile.py: the weight-stationary kernel is bound by the bytes it moves through L2 -- ~4.6 KB/clk for the whole
// chip, loads of A plus stores of C, whatever the mix (no stores: -1077 cycles per tile; no A loads: -380; neither: 2745 of a
// 2304-cycle MMA floor) -- i.e. by 768/BN + 2 bytes per output element, and BN is capped by the shared memory the resident
// block needs (192 columns = 144 KB).  A pair of SMs running tcgen05.mma.cta_group::2 shares ONE B operand: each CTA keeps 128
// of the pair's 256 columns resident (96 KB), loads its own 128 rows of A, and the leader's M = 256 instructions read both
// halves.  5 bytes per element instead of 6 (BN = 192) or 8 (BN = 128, the CRF head), half the operand bytes read from each
// SM's shared memory per MMA, and a 6-stage ring.  Roles per CTA as above; only rank 0 issues MMAs, TMA loads of both CTAs
// complete on rank 0's `full` barrier, tcgen05.commit multicasts `empty` / `tfull` to both CTAs, the epilogue warps of both
// arrive on rank 0's `tempty`.
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
            "r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_mma_ss_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {     // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}

struct PairSmem {
    static constexpr int kStages = 6;
    static constexpr uint32_t kBres = 0;                                  // [WS_KB][128 rows][128 B]: this CTA's half of B
    static constexpr uint32_t kRing = WS_KB * 128 * 128;                  // kStages x (128 rows x 128 B) of this CTA's A rows
    static constexpr uint32_t kEpi = kRing + kStages * BM * BK * 2;
    static constexpr uint32_t kBias = kEpi + EPI_WARPS * 2048;
    static constexpr uint32_t kBars = kBias + EPI_WARPS * 512;            // 2 * kStages + 6 mbarriers
    static constexpr uint32_t kTileQ = kBars + 8 * (2 * kStages + 6);
    static constexpr uint32_t kTotal = kTileQ + 12 * TQ + 64 + 1024;
};
static_assert(PairSmem::kTotal <= SMEM_LIMIT, "pair kernel: shared memory");

// A pair is bound to ONE 256-column block for its whole life and draws 256-row blocks from that block's ticket counter, like
// the single-CTA kernel: the pairs of all column blocks then sweep over A at the same pace and every A tile comes from HBM
// once and from L2 for the other column blocks.  (One global counter over column-block-major tiles keeps all 74 pairs busy
// for any N, but re-reads the 655 MB of A from HBM once per column block: measured 806 vs 989 TFLOP/s.)
// RES = 1: weight-stationary (above).  RES = 0: both operands streamed -- a stage holds this CTA's 128 rows of A and its 128
// of the tile's 256 rows of B (32 KB, 6 stages), tiles come from ONE ticket counter with the column block fastest (the pairs
// work on a few row blocks of A at a time, across all column blocks: A and the weights are re-read from L2), any K.  Used by
// the transformer's fc2 (K = 2048) instead of the single-CTA streaming kernel.
template <int RES>
__global__ void __launch_bounds__(THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 __half* __restrict__ C, long long ldc, int M, int N, int K, GemmEpilogue ep, int* __restrict__ ctr) {
    using S = PairSmem;
    constexpr int ST = S::kStages, BNP = 256;
    constexpr uint32_t STAGE = RES ? BM * BK * 2 : 2 * BM * BK * 2;      // bytes per ring stage of one CTA
    constexpr uint32_t RING0 = RES ? S::kRing : 0;                        // the ring takes the resident block's place when streaming
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + S::kBars;
    auto full_bar = [&](int s) { return bars + 8u * s; };                 // used in rank 0 only
    auto empty_bar = [&](int s) { return bars + 8u * (ST + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * ST + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * ST + 2 + s); }; // used in rank 0 only
    const uint32_t bres_bar = bars + 8u * (2 * ST + 4);                    // used in rank 0 only
    const uint32_t tmem_slot = bars + 8u * (2 * ST + 5);
    unsigned char* gen_base = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t tq_bars = base + S::kTileQ, tq_tiles = base + S::kTileQ + 8 * TQ;
    volatile int* tq_tiles_gen = reinterpret_cast<volatile int*>(gen_base + S::kTileQ + 8 * TQ);
    auto take = [&](int q) {
        mbar_wait_cluster(tq_bars + 8u * (uint32_t)(q & (TQ - 1)), (uint32_t)((q / TQ) & 1));
        return tq_tiles_gen[q & (TQ - 1)];
    };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int m_blocks = (M + 2 * BM - 1) / (2 * BM), n_blocks = N / BNP;   // 256-row blocks, 256-column blocks
    const int k_blocks = (K + BK - 1) / BK;
    const int nb_fixed = (int)(blockIdx.x / 2) % n_blocks;    // RES: the pair's column block
    const int tickets = RES ? m_blocks : m_blocks * n_blocks;
    int* my_ctr = RES ? ctr + nb_fixed : ctr;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_a));
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_b));
        for (int s = 0; s < ST; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 2 * EPI_WARPS);   // the epilogue warps of both CTAs
        }
        mbar_init(bres_bar, 1);
        for (int s = 0; s < TQ; ++s) mbar_init(tq_bars + 8u * s, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tmem_slot), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + S::kBars + 8u * (2 * ST + 5));

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (elect_one_sync()) {
            auto post = [&](int q, int tile) {
                const uint32_t slot = (uint32_t)(q & (TQ - 1));
                for (uint32_t r = 0; r < 2u; ++r) {
                    st_cluster_s32(mapa(tq_tiles + 4u * slot, r), tile);
                    mbar_arrive_remote(mapa(tq_bars + 8u * slot, r));
                }
            };
            const uint32_t l_bres = mapa(bres_bar, 0);
            int stage = 0;
            uint32_t phase = 0;
            int next = 0;
            if (rank == 0) {
                const int first = atomicAdd(my_ctr, 1);
                post(0, first < tickets ? first : -1);
            }
            for (int q = 0;; ++q) {
                const int tile = take(q);
                if (tile < 0) break;
                if (rank == 0) next = atomicAdd(my_ctr, 1);
                const int mb = RES ? tile : tile / n_blocks, nb = RES ? nb_fixed : tile - mb * n_blocks;
                if (RES && q == 0) {
                    if (rank == 0) mbar_expect_tx(bres_bar, 2u * (uint32_t)k_blocks * 128 * 128);
                    for (int kb = 0; kb < k_blocks; ++kb)
                        tma_load_2d_pair(base + S::kBres + kb * (128 * 128), &map_b, l_bres, kb * BK, nb * BNP + (int)rank * 128);
                }
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t l_full = mapa(full_bar(stage), 0);
                    if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * STAGE);
                    tma_load_2d_pair(base + RING0 + stage * STAGE, &map_a, l_full, kb * BK, mb * 2 * BM + (int)rank * BM);
                    if (!RES)
                        tma_load_2d_pair(base + RING0 + stage * STAGE + BM * BK * 2, &map_b, l_full, kb * BK, nb * BNP + (int)rank * 128);
                    if (++stage == ST) { stage = 0; phase ^= 1; }
                }
                if (rank == 0) post(q + 1, next < tickets ? next : -1);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: rank 0 only, M = 256 across the pair =====
        if (rank == 0 && elect_one_sync()) {
            const uint32_t idesc = tc_idesc_f16(2 * BM, BNP);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int q = 0; take(q) >= 0; ++q) {
                if (RES && q == 0) mbar_wait_cluster(bres_bar, 0);      // both halves of the column block have landed
                mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BNP);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait_cluster(full_bar(stage), phase);
                    tc_fence_after();
                    const uint64_t adesc = tc_smem_desc_sw128(base + RING0 + stage * STAGE);
                    const uint64_t bdesc = tc_smem_desc_sw128(RES ? base + S::kBres + kb * (128 * 128)
                                                                  : base + RING0 + stage * STAGE + BM * BK * 2);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        tc_mma_ss_pair(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit_pair(empty_bar(stage));
                    if (++stage == ST) { stage = 0; phase ^= 1; }
                }
                tc_commit_pair(tfull_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps (both CTAs): this CTA's 128 rows of the pair's tile =====
        const int quarter = warp & 3, ew = warp - 2, set = ew >> 2;
        unsigned char* tbuf = gen_base + S::kEpi + ew * 2048;
        __half* sbias = reinterpret_cast<__half*>(gen_base + S::kBias + ew * 512);
        const uint32_t l_tempty0 = mapa(tempty_bar(0), 0), l_tempty1 = mapa(tempty_bar(1), 0);
        int acc = 0, cur_nb = -1;
        uint32_t acc_phase = 0;
        for (int q = 0;; ++q) {
            const int tile = take(q);
            if (tile < 0) break;
            const int mb = RES ? tile : tile / n_blocks, nb = RES ? nb_fixed : tile - mb * n_blocks;
            if (nb != cur_nb) {      // before the wait: its latency hides behind the mainloop
                stage_bias<BNP>(sbias, ep.bias, nb, N, lane);
                cur_nb = nb;
            }
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            epilogue_tile<BNP, 0>(tmem_base + (uint32_t)(acc * BNP), tbuf, sbias, C, ldc, M, N, mb * 2 + (int)rank, nb, quarter, set,
                                  lane, ep);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(acc ? l_tempty1 : l_tempty0);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // both CTAs are done with the pair's tensor memory and barriers
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u));
    if (threadIdx.x == 0) release_tile_counters(ctr, RES ? n_blocks : 1);
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
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

// 2-D fp16 (or int8) tensor [rows][cols] with row stride `ld` elements, box = box_rows x 128 bytes, 128-B swizzle.
int make_map(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows, bool i8 = false) {
    EncodeTiledFn fn = get_encode_fn();
    B200_REQUIRE(fn != nullptr, "gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * (i8 ? 1 : 2)};
    cuuint32_t box[2] = {(cuuint32_t)(i8 ? 2 * BK : BK), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, i8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
                 rows, cols, ld);
    return 0;
}

// Pool of zero-initialised ticket counters (per device).  Eager launches take theirs round-robin from the first half
// (a region is reused only ~100k launches later, long after its launch has reset it); launches recorded into a CUDA graph
// take theirs from the second half and keep them for the life of the process, because a replay uses the same addresses
// every time and must never share them with a concurrent eager launch.
constexpr int POOL_INTS = 1 << 21;   // 8 MB
struct CounterPool {
    int* base = nullptr;
    int eager = 0, pinned = POOL_INTS / 2;
};
static CounterPool g_pools[16];
static std::mutex g_pool_mutex;

int take_counters(int n, cudaStream_t stream, int** out) {
    int dev = 0;
    B200_CHECK_CUDA(cudaGetDevice(&dev));
    B200_REQUIRE(dev >= 0 && dev < 16, "gemm_tc: device ordinal %d is not supported", dev);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    B200_CHECK_CUDA(cudaStreamIsCapturing(stream, &cap));
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    CounterPool& p = g_pools[dev];
    if (p.base == nullptr) {
        B200_REQUIRE(cap == cudaStreamCaptureStatusNone,
                     "gemm_tc: the first GEMM on a device must not run inside a CUDA graph capture (run one warm-up step)");
        B200_CHECK_CUDA(cudaMalloc(&p.base, sizeof(int) * POOL_INTS));
        B200_CHECK_CUDA(cudaMemset(p.base, 0, sizeof(int) * POOL_INTS));
    }
    if (cap != cudaStreamCaptureStatusNone) {
        B200_REQUIRE(p.pinned + n <= POOL_INTS, "gemm_tc: out of ticket counters for graph-captured launches");
        *out = p.base + p.pinned;
        p.pinned += n;
    } else {
        if (p.eager + n > POOL_INTS / 2) p.eager = 0;
        *out = p.base + p.eager;
        p.eager += n;
    }
    return 0;
}

template <int BN>
int launch_tc(const __half* A, long long lda, const __half* B, __half* C, long long ldc, int M, int N, int K,
              const GemmEpilogue& ep, int max_ctas, cudaStream_t stream) {
    CUtensorMap map_a, map_b;
    int rc = make_map(&map_a, A, M, K, lda, BM);
    if (rc) return rc;
    rc = make_map(&map_b, B, N, K, K, BN);
    if (rc) return rc;
    auto kern = gemm_tc_kernel<BN>;
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcSmem<BN>::kTotal));
    int dev = 0, sms = 0;
    B200_CHECK_CUDA(cudaGetDevice(&dev));
    B200_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int grid = tiles < sms ? tiles : sms;
    int* ctr = nullptr;
    rc = take_counters(2, stream, &ctr);
    if (rc) return rc;
    kern<<<grid, THREADS, TcSmem<BN>::kTotal, stream>>>(map_a, map_b, C, ldc, M, N, K, ep, ctr);
    B200_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <int BN, int I8 = 0>
int launch_ws(const void* A, long long lda, const void* B, __half* C, long long ldc, int M, int N, int K,
              const GemmEpilogue& ep, int max_ctas, cudaStream_t stream, const float* col_scale = nullptr) {
    CUtensorMap map_a, map_b;
    int rc = make_map(&map_a, A, M, K, lda, BM, I8);
    if (rc) return rc;
    rc = make_map(&map_b, B, N, K, K, BN, I8);
    if (rc) return rc;
    auto kern = gemm_ws_kernel<BN, I8>;
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WsSmem<BN, I8>::kTotal));
    int dev = 0, sms = 0;
    B200_CHECK_CUDA(cudaGetDevice(&dev));
    B200_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
    const int m_blocks = (M + BM - 1) / BM, n_blocks = (N + BN - 1) / BN;
    int per_block = sms / n_blocks;            // CTAs bound to one column block
    if (per_block > m_blocks) per_block = m_blocks;
    if (per_block < 1) per_block = 1;
    int* ctr = nullptr;
    rc = take_counters(n_blocks + 1, stream, &ctr);
    if (rc) return rc;
    kern<<<per_block * n_blocks, THREADS, WsSmem<BN, I8>::kTotal, stream>>>(map_a, map_b, C, ldc, M, N, K, ep, ctr, col_scale, gemm_debug());
    B200_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int pair_enabled() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("B200_GEMM_PAIR");
        mode = e ? atoi(e) : PAIR_DEFAULT;
    }
    return mode;
}

template <int RES>
int launch_pair(const __half* A, long long lda, const __half* B, __half* C, long long ldc, int M, int N, int K,
                const GemmEpilogue& ep, int max_ctas, cudaStream_t stream) {
    CUtensorMap map_a, map_b;
    int rc = make_map(&map_a, A, M, K, lda, BM);
    if (rc) return rc;
    rc = make_map(&map_b, B, N, K, K, 128);
    if (rc) return rc;
    auto kern = gemm_pair_kernel<RES>;
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PairSmem::kTotal));
    int dev = 0, sms = 0;
    B200_CHECK_CUDA(cudaGetDevice(&dev));
    B200_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
    const int m_blocks = (M + 2 * BM - 1) / (2 * BM), n_blocks = N / 256;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = PairSmem::kTotal;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static int s_max_pairs = 0;
    if (s_max_pairs == 0) {
        cfg.gridDim = dim3(2 * 128);
        int n = 0;
        B200_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
        s_max_pairs = n > 0 ? n : 1;
    }
    int pairs = sms / 2 < s_max_pairs ? sms / 2 : s_max_pairs;
    int n_ctr = 1;
    if (RES) {          // pairs bound to column blocks, one ticket counter per block
        int per_block = pairs / n_blocks;
        if (per_block > m_blocks) per_block = m_blocks;
        if (per_block < 1) per_block = 1;
        pairs = per_block * n_blocks;
        n_ctr = n_blocks;
    } else if (pairs > m_blocks * n_blocks) {
        pairs = m_blocks * n_blocks;
    }
    cfg.gridDim = dim3(pairs * 2);
    int* ctr = nullptr;
    rc = take_counters(n_ctr + 1, stream, &ctr);
    if (rc) return rc;
    B200_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, map_a, map_b, C, ldc, M, N, K, ep, ctr));
    return 0;
}

}  // namespace

int launch_gemm_tc(const __half* A, long long lda, const __half* B, __half* C, long long ldc, int M, int N, int K,
                   const GemmEpilogue& ep, int max_ctas, cudaStream_t stream, bool force_pair) {
    B200_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0 && ((uintptr_t)C % 16) == 0,
                 "gemm_tc: operands must be 16-byte aligned");
    const char* env = getenv("B200_GEMM_WS");
    const bool ws_ok = !(env && env[0] == '0') && K <= WS_KB * BK && M >= 4 * BM;
    // cta_group::2 pairs: faster in isolation (input projection 926-966 -> 989 TFLOP/s, N = 4096 head 725-748 -> 875, sup fc2
    // 6.81 -> 6.12 ms), but not the default: the pipelined step gains nothing from the faster GEMM (17.3 vs 17.4 ms: it is
    // bound by SM-milliseconds), and a pair needs both SMs of a TPC free at once while the other batch's
    // single-CTA kernels come and go.
    if ((force_pair || pair_enabled()) && N % 256 == 0 && (force_pair || M >= 64 * BM)) {
        if (K <= WS_KB * BK && N / 256 <= 64 && ep.act != B200_ACT_SWIGLU)
            return launch_pair<1>(A, lda, B, C, ldc, M, N, K, ep, max_ctas, stream);
        if (K >= 1024 || force_pair) return launch_pair<0>(A, lda, B, C, ldc, M, N, K, ep, max_ctas, stream);
    }
    B200_REQUIRE(!force_pair, "gemm: the pair kernels need N (%d) to be a multiple of 256", N);
    if (ws_ok && N % 192 == 0 && N / 192 <= 64) return launch_ws<192, 0>(A, lda, B, C, ldc, M, N, K, ep, max_ctas, stream);
    if (ws_ok && N % 128 == 0 && N / 128 <= 64) return launch_ws<128, 0>(A, lda, B, C, ldc, M, N, K, ep, max_ctas, stream);
    if (N % 256 == 0) return launch_tc<256>(A, lda, B, C, ldc, M, N, K, ep, max_ctas, stream);
    return launch_tc<128>(A, lda, B, C, ldc, M, N, K, ep, max_ctas, stream);
}

// C = act(scale[col] * (A_i8 B_i8^T) + bias): int8 operands (A [M][K] row stride lda bytes, B [N][K]), s32 accumulation on
// kind::i8 tensor cores, weight-stationary kernel only (K <= 768, N a multiple of 192 or 128).
int launch_gemm_i8(const int8_t* A, long long lda, const int8_t* B, const float* col_scale, __half* C, long long ldc, int M,
                   int N, int K, const GemmEpilogue& ep, int max_ctas, cudaStream_t stream) {
    B200_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0 && ((uintptr_t)C % 16) == 0 && col_scale != nullptr,
                 "gemm_i8: operands must be 16-byte aligned and a column scale is required");
    B200_REQUIRE(K % 16 == 0 && lda % 16 == 0 && K <= WS_KB * BK && ep.act != B200_ACT_SWIGLU,
                 "gemm_i8: K (%d) and lda (%lld) must be multiples of 16 bytes, K <= %d", K, lda, WS_KB * BK);
    if (N % 192 == 0 && N / 192 <= 64) return launch_ws<192, 1>(A, lda, B, C, ldc, M, N, K, ep, max_ctas, stream, col_scale);
    B200_REQUIRE(N % 128 == 0 && N / 128 <= 64, "gemm_i8: N (%d) must be a multiple of 192 or 128", N);
    return launch_ws<128, 1>(A, lda, B, C, ldc, M, N, K, ep, max_ctas, stream, col_scale);
}

// copies the per-CTA cycle counters of the last weight-stationary launch (B200_GEMM_DEBUG=1): 160 x 8 values
int copy_gemm_profile(long long* host_out) {
    B200_CHECK_CUDA(cudaDeviceSynchronize());
    B200_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_gemm_prof, sizeof(long long) * 160 * 8));
    return 0;
}
