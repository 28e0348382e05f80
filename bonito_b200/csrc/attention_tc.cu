// Windowed softmax attention on the 5th-generation tensor cores (sup v5: head_dim 64, window 127 / 128).
// Reference semantics: bonito/transformer/model.py:42-79 -- flash_attn_qkvpacked_func(window_size=(wl, wr)), non-causal,
// softmax scale 1/sqrt(head_dim); the rotary embedding has been applied to q and k in place before (rotary_kernel).
//
// One CTA per (chunk, head, 128-query tile).  With wl, wr <= 128 the queries of a tile see at most three 128-key tiles
// (the one before, its own, the one after), so the whole band of scores fits in tensor memory at once:
//   * TMA (3-D tensor map over qkv [N][T][3*heads*64], SWIZZLE_128B, out-of-range rows zero-filled) brings Q [128 x 64],
//     K_j [128 x 64] and V_j [128 keys x 64] (j = 0..2) into shared memory;
//   * S_j = Q K_j^T: 4 tcgen05.mma (M=128, N=128, K=16) per key tile, fp32 accumulators in TMEM columns [128j, 128j+128);
//   * four softmax warps, one thread per query row (tcgen05.ld 32x32b: thread i of warp w owns TMEM lane 32w+i): each key
//     tile is normalised on its OWN row maximum m_j -- the 128 scores of a row and tile live in registers between the
//     maximum and the exponentials, so S is read from TMEM exactly once -- and P_j = 2^((s - m_j) * scale) is written back
//     as fp16 pairs into the first 64 columns of S_j (tcgen05.st), i.e. as a TMEM-resident A operand;
//   * O_j = P_j V_j: 8 tcgen05.mma (M=128, N=64, K=16) per key tile with A from TMEM and B = V_j straight from its TMA
//     layout (rows = keys = K, 128-byte rows of 64 head dims: the MN-major SWIZZLE_128B operand, "transpose B" bit of the
//     instruction descriptor), accumulated into the LAST 64 columns of S_j; PV_j runs while the softmax warps work on j+1;
//   * epilogue: O = sum_j O_j 2^((m_j - M) scale) / sum_j l_j 2^((m_j - M) scale), M = max_j m_j -- three independent
//     partial softmaxes combined per row, no running-maximum rescaling of accumulators in flight.
// Masked scores (outside the window or outside the chunk) get P = 0; 32-column pieces that are masked for a whole warp are
// neither read nor exponentiated.
#include <cuda.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace {

constexpr int BQ = 128, BKV = 128, HD = 64, NKT = 3;
constexpr int THREADS = 160;                       // 4 softmax warps + 1 TMA / MMA warp
constexpr uint32_t TILE_BYTES = BQ * HD * 2;       // 16 KB: one 128 x 64 fp16 tile (128-byte rows)
constexpr uint32_t OFF_Q = 0, OFF_K = TILE_BYTES, OFF_V = OFF_K + NKT * TILE_BYTES, OFF_BARS = OFF_V + NKT * TILE_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_BARS + 128 + 1024;
constexpr uint32_t TMEM_COLS = 512;                // S_j | P_j | O_j share columns [128j, 128j+128), j < 3

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::
            "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, __half* __restrict__ out, int T, int NH, int wl, int wr,
                    float scale_log2e) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024-B alignment
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + OFF_BARS;
    const uint32_t bar_qk = bars, bar_v = bars + 8, bar_o = bars + 16;
    auto bar_s = [&](int j) { return bars + 24u + 8u * (uint32_t)j; };
    auto bar_p = [&](int j) { return bars + 48u + 8u * (uint32_t)j; };
    const uint32_t tmem_slot = bars + 72;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y, n = blockIdx.z;
    const int k0 = q0 - BKV;                       // first key of key tile 0

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_qkv));
        mbar_init(bar_qk, 1);
        mbar_init(bar_v, 1);
        mbar_init(bar_o, 1);
        for (int j = 0; j < NKT; ++j) {
            mbar_init(bar_s(j), 1);
            mbar_init(bar_p(j), 4);                // one arrive per softmax warp
        }
        mbar_fence_init();
    }
    if (warp == 4) tc_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + OFF_BARS + 72);

    if (warp == 4) {
        // ===== TMA producer + MMA issuer =====
        if (elect_one_sync()) {
            const int cq = h * HD, ck = NH * HD + h * HD, cv = 2 * NH * HD + h * HD;
            mbar_expect_tx(bar_qk, (1 + NKT) * TILE_BYTES);
            tma_load_3d(base + OFF_Q, &map_qkv, bar_qk, cq, q0, n);
            for (int j = 0; j < NKT; ++j) tma_load_3d(base + OFF_K + j * TILE_BYTES, &map_qkv, bar_qk, ck, k0 + j * BKV, n);
            mbar_expect_tx(bar_v, NKT * TILE_BYTES);
            for (int j = 0; j < NKT; ++j) tma_load_3d(base + OFF_V + j * TILE_BYTES, &map_qkv, bar_v, cv, k0 + j * BKV, n);

            // S_j = Q K_j^T  (both operands K-major, 128-byte swizzled rows of 64 head dims)
            mbar_wait(bar_qk, 0);
            tc_fence_after();
            constexpr uint32_t idesc_s = tc_idesc_f16(BQ, BKV);
            const uint64_t qdesc = tc_smem_desc_sw128(base + OFF_Q);
#pragma unroll
            for (int j = 0; j < NKT; ++j) {
                const uint64_t kdesc = tc_smem_desc_sw128(base + OFF_K + j * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    tc_mma_ss(tmem_base + (uint32_t)(j * 128), qdesc + 2u * k, kdesc + 2u * k, idesc_s, k != 0 ? 1u : 0u);
                tc_commit(bar_s(j));
            }
            // O_j = P_j V_j  (A = P_j from TMEM: lane = query, column c = keys 2c, 2c+1; B = V_j MN-major: K = keys are the
            // 128-byte rows, 8 keys per 1024-byte swizzle atom, so one K = 16 step advances the descriptor by 2048 bytes)
            mbar_wait(bar_v, 0);
            constexpr uint32_t idesc_o = tc_idesc_f16(BQ, HD) | (1u << 16);   // bit 16: B is MN-major
#pragma unroll
            for (int j = 0; j < NKT; ++j) {
                mbar_wait(bar_p(j), 0);
                tc_fence_after();
                const uint64_t vdesc = tc_smem_desc_sw128(base + OFF_V + j * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k)
                    tc_mma_ts(tmem_base + (uint32_t)(j * 128 + 64), tmem_base + (uint32_t)(j * 128 + 8 * k),
                              vdesc + (uint64_t)(k * (2048 >> 4)), idesc_o, k != 0 ? 1u : 0u);
            }
            tc_commit(bar_o);
        }
        __syncwarp();
    } else {
        // ===== softmax warps: thread = query row =====
        const int r = warp * 32 + lane, q = q0 + r;
        // visible band of this row in band columns c = key - k0 (0 .. 383)
        const int lo = max(BKV + r - wl, -k0), hi = min(BKV + r + wr, T - 1 - k0);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        float m[NKT], l[NKT];
#pragma unroll
        for (int j = 0; j < NKT; ++j) {
            const int a = max(lo - j * BKV, 0), b = min(hi - j * BKV, BKV - 1);    // visible columns of tile j: [a, b]
            mbar_wait(bar_s(j), 0);
            tc_fence_after();
            uint32_t s[BKV];
            // pieces of 32 columns; a piece that no row of this warp sees is not even read
            bool need[4];
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
                need[pc] = __any_sync(0xffffffffu, a <= pc * 32 + 31 && b >= pc * 32);
                if (need[pc]) tc_ld_32x32b_x32(lane_addr + (uint32_t)(j * 128 + pc * 32), *reinterpret_cast<uint32_t(*)[32]>(&s[pc * 32]));
            }
            tc_wait_ld();
            float mx = -1e30f;
#pragma unroll
            for (int pc = 0; pc < 4; ++pc)
                if (need[pc]) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int col = pc * 32 + c;
                        if (col >= a && col <= b) mx = fmaxf(mx, __uint_as_float(s[col]));
                    }
                }
            m[j] = mx;
            const float mb = mx * scale_log2e;
            float sum = 0.f;
#pragma unroll
            for (int half = 0; half < 2; ++half) {          // 64 scores -> 32 fp16 pairs -> P columns [32*half, +32)
                uint32_t pk[32];
#pragma unroll
                for (int c2 = 0; c2 < 32; ++c2) {
                    const int col = half * 64 + 2 * c2, pc = col >> 5;
                    float p0 = 0.f, p1 = 0.f;
                    if (need[pc]) {
                        if (col >= a && col <= b) p0 = ex2_approx(fmaf(__uint_as_float(s[col]), scale_log2e, -mb));
                        if (col + 1 >= a && col + 1 <= b) p1 = ex2_approx(fmaf(__uint_as_float(s[col + 1]), scale_log2e, -mb));
                    }
                    sum += p0 + p1;
                    const __half2 h2 = __floats2half2_rn(p0, p1);
                    pk[c2] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                tc_st_32x32b_x32(lane_addr + (uint32_t)(j * 128 + half * 32), pk);
            }
            l[j] = sum;
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p(j));
        }
        // ===== epilogue: combine the three partial softmaxes of the row =====
        const float M = fmaxf(fmaxf(m[0], m[1]), m[2]);
        float f[NKT], L = 0.f;
#pragma unroll
        for (int j = 0; j < NKT; ++j) {
            f[j] = ex2_approx(fmaxf((m[j] - M) * scale_log2e, -126.f));
            if (m[j] <= -1e30f) f[j] = 0.f;
            L = fmaf(l[j], f[j], L);
        }
        const float inv = L > 0.f ? 1.0f / L : 0.f;
        mbar_wait(bar_o, 0);
        tc_fence_after();
        __half* dst = out + ((size_t)n * T + q) * (size_t)(NH * HD) + h * HD;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t o0[32], o1[32], o2[32];
            tc_ld_32x32b_x32(lane_addr + (uint32_t)(0 * 128 + 64 + half * 32), o0);
            tc_ld_32x32b_x32(lane_addr + (uint32_t)(1 * 128 + 64 + half * 32), o1);
            tc_ld_32x32b_x32(lane_addr + (uint32_t)(2 * 128 + 64 + half * 32), o2);
            tc_wait_ld();
            if (q < T) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    __half2 hh[4];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const int c = g * 8 + 2 * p;
                        const float v0 = (__uint_as_float(o0[c]) * f[0] + __uint_as_float(o1[c]) * f[1] + __uint_as_float(o2[c]) * f[2]) * inv;
                        const float v1 = (__uint_as_float(o0[c + 1]) * f[0] + __uint_as_float(o1[c + 1]) * f[1] + __uint_as_float(o2[c + 1]) * f[2]) * inv;
                        hh[p] = __floats2half2_rn(v0, v1);
                    }
                    *reinterpret_cast<uint4*>(dst + half * 32 + g * 8) = *reinterpret_cast<const uint4*>(hh);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) tc_dealloc(tmem_base, TMEM_COLS);
}

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
    dim3 grid((T + BQ - 1) / BQ, NH, N);
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)HD);
    attention_tc_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(map, out, T, NH, wl, wr, scale_log2e);
    B200_CHECK_CUDA(cudaGetLastError());
    return 0;
}
