// Micro-benchmarks behind b200_debug_mma_bench: issue-to-completion cycles of a chain of tcgen05.mma
// (M=128, K=16, fp16) for different N and A sources.  Timing aid only.
#include "tc_common.cuh"

namespace {

__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int ts_mode, int n, int iters, int chains, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* g = smem_raw + (base - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(g)[i] = 0x3c003c00u;  // 1.0h
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    if (warp == 0) tc_alloc(smem_u32(&slot), 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    {   // A in TMEM columns [256, 288)
        uint32_t v[32];
        for (int c = 0; c < 32; ++c) v[c] = 0x3c003c00u;
        tc_st_32x32b_x32(tb + ((uint32_t)(warp * 32) << 16) + 448, v);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0 && elect_one_sync()) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t adesc = tc_smem_desc_sw128(base);
        const uint64_t bdesc = tc_smem_desc_sw128(base + 16384);
        unsigned long long g0, g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        const long long t0 = clock64();
        int c = 0;
        for (int i = 0; i < iters; ++i) {
            const uint32_t k = i & 3;
            const uint32_t d = tb + (uint32_t)(c * n);
            if (ts_mode) tc_mma_ts(d, tb + 448 + k * 8, bdesc + 2u * k, idesc, i >= chains ? 1u : 0u);
            else tc_mma_ss(d, adesc + 2u * k, bdesc + 2u * k, idesc, i >= chains ? 1u : 0u);
            if (++c == chains) c = 0;
        }
        const long long t1 = clock64();
        tc_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = (long long)(g1 - g0); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tb, 512);
}

// ---- h all-gather micro-benchmark -------------------------------------------------------------------------------
// The communication skeleton of the tile recurrent kernel without the math: a cluster of 6 CTAs, per CTA 3 sub-tiles x 8
// sender warps; every step each sender warp delivers its 256-byte block into the h tile (sub, parity) of ALL six CTAs; a
// consumer warp waits for the complete tile (12288 B of transaction bytes), re-arms the barrier and releases the sender
// warps of that sub-tile, which spin `delay` cycles (the MMA + cell update of the real kernel) and send the next block.
//   mode 0  one cp.async.bulk shared::cta -> shared::cluster per peer and warp (256 B)            [DSMEM]
//   mode 2  block -> global staging (L2), fence.proxy.async.global, ONE multicast cp.async.bulk per warp (256 B)
//   mode 3  as 2, but the 8 warps of a sub-tile meet at a named barrier and one of them multicasts their 2 KB
//   mode 4  DSMEM with one 2 KB bulk copy per (sub-tile, peer) behind a named barrier
// out[0] = cycles CTA 0 spent in the loop, out[1] = steps.
namespace xb {
constexpr int CS = 6, NS = 3, EW = 8, SN = 16;
constexpr uint32_t HT = 48 * SN * 16;           // 12288
constexpr uint32_t OFF_H = 0, OFF_STAGE = NS * 2 * HT, OFF_BARS = OFF_STAGE + NS * EW * 256 * 2;
constexpr uint32_t SMEM = OFF_BARS + 256 + 1024;
constexpr int THREADS = (NS * EW + 1) * 32;

__device__ __forceinline__ void bulk_multicast(uint32_t dst, const void* gsrc, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n" ::
            "r"(dst), "l"(gsrc), "r"(bytes), "r"(bar), "h"(mask)
        : "memory");
}

__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(THREADS, 1)
exchange_bench_kernel(int mode, int steps, int delay, unsigned char* __restrict__ staging, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + OFF_BARS;
    auto hfull = [&](int sub, int p) { return bars + 8u * (uint32_t)(sub * 2 + p); };
    auto go = [&](int sub) { return bars + 8u * (uint32_t)(6 + sub); };
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    const int cluster = blockIdx.x / CS;
    // global staging: [cluster][parity][sub][48 k-chunks][256 B]
    unsigned char* stg = staging + (size_t)cluster * 2 * NS * HT;
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(bars + 8 * i, 1);
        for (int i = 0; i < 3; ++i) mbar_init(bars + 8 * (6 + i), 1);
        mbar_fence_init();
        for (int sub = 0; sub < NS; ++sub) {
            if (steps > 0) mbar_expect_tx(hfull(sub, 0), HT);
            if (steps > 1) mbar_expect_tx(hfull(sub, 1), HT);
        }
    }
    for (int i = tid; i < (int)(OFF_BARS / 16); i += THREADS) reinterpret_cast<uint4*>(gbase)[i] = make_uint4(tid, i, 0, 0);
    fence_proxy_async();
    __syncthreads();
    cluster_sync_all();
    long long t0 = 0;
    if (warp == NS * EW) {
        // consumer: tile complete -> re-arm -> release the senders of that sub-tile
        t0 = clock64();
        for (int s = 0; s < steps; ++s) {
            const int p = s & 1;
            for (int sub = 0; sub < NS; ++sub) {
                mbar_wait(hfull(sub, p), (uint32_t)((s >> 1) & 1));
                if (lane == 0) {
                    if (s + 2 < steps) mbar_expect_tx(hfull(sub, p), HT);
                    mbar_arrive(go(sub));
                }
                __syncwarp();
            }
        }
        if (lane == 0 && blockIdx.x == 0) { out[0] = clock64() - t0; out[1] = steps; }
    } else {
        const int sub = warp / EW, ew = warp % EW;
        const uint32_t kchunk = rank * EW + ew;                                  // this warp's k-chunk of the tile
        const uint32_t stage_off = OFF_STAGE + (uint32_t)(sub * EW + ew) * 512;   // two parities of 256 B
        uint32_t peer_shift[CS];
#pragma unroll
        for (int d = 0; d < CS; ++d) peer_shift[d] = mapa(base, (rank + 1u + (uint32_t)d) % CS) - base;
        for (int s = 0; s < steps; ++s) {
            const int p = s & 1;
            if (s > 0) mbar_wait(go(sub), (uint32_t)((s - 1) & 1));   // step s-1 of this sub-tile was complete everywhere
            const long long t = clock64();
            while (clock64() - t < delay) {}
            const uint32_t dst = base + OFF_H + (uint32_t)(sub * 2 + p) * HT;
            const uint32_t src = base + stage_off + p * 256;
            if (mode == 0) {
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one_sync()) {
#pragma unroll
                    for (int d = 0; d < CS; ++d)
                        bulk_copy_to_peer(dst + kchunk * 256 + peer_shift[d], src, 256, hfull(sub, p) + peer_shift[d]);
                }
            } else if (mode == 4) {
                fence_proxy_async_smem();
                asm volatile("bar.sync %0, 256;\n" ::"r"(1 + sub) : "memory");
                if (ew == 0 && elect_one_sync()) {
                    const uint32_t src8 = base + OFF_STAGE + (uint32_t)(sub * EW) * 512 + 0;   // (contiguity is not needed for timing)
#pragma unroll
                    for (int d = 0; d < CS; ++d)
                        bulk_copy_to_peer(dst + rank * 2048 + peer_shift[d], src8, 2048, hfull(sub, p) + peer_shift[d]);
                }
            } else {
                unsigned char* g = stg + (size_t)(p * NS + sub) * HT + kchunk * 256;
                if (lane < 16) reinterpret_cast<uint4*>(g)[lane] = reinterpret_cast<const uint4*>(gbase + stage_off + p * 256)[lane];
                asm volatile("fence.proxy.async.global;\n" ::: "memory");
                if (mode == 2) {
                    __syncwarp();
                    if (elect_one_sync()) bulk_multicast(dst + kchunk * 256, g, 256, hfull(sub, p), (uint16_t)0x3f);
                } else {
                    asm volatile("bar.sync %0, 256;\n" ::"r"(1 + sub) : "memory");
                    if (ew == 0 && elect_one_sync())
                        bulk_multicast(dst + rank * 2048, stg + (size_t)(p * NS + sub) * HT + rank * 2048, 2048, hfull(sub, p),
                                       (uint16_t)0x3f);
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    cluster_sync_all();
}
}  // namespace xb

// dummy cluster kernel for occupancy queries
__global__ void cluster_probe_kernel(int* out) {
    extern __shared__ unsigned char smem_raw[];
    if (threadIdx.x == 0 && out != nullptr) atomicAdd(out, (int)cluster_ctarank() + (smem_raw[0] & 0));
}

}  // namespace

int launch_exchange_bench(int mode, int steps, int delay, int clusters, unsigned char* staging, long long* out,
                          cudaStream_t stream) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(xb::exchange_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xb::SMEM + 100 * 1024));
    xb::exchange_bench_kernel<<<clusters * xb::CS, xb::THREADS, xb::SMEM + 100 * 1024, stream>>>(mode, steps, delay, staging, out);
    B200_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// how many clusters of `cluster_size` CTAs (threads, dynamic shared memory as given) the device holds at once
int debug_max_clusters(int cluster_size, int threads, int smem_bytes) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster_size * 64);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaFuncSetAttribute(cluster_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    if (cluster_size > 8 &&
        cudaFuncSetAttribute(cluster_probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, cluster_probe_kernel, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return n;
}

int launch_mma_bench(int ts_mode, int n, int iters, int chains, int blocks, long long* out, cudaStream_t stream) {
    const int smem = 16384 + 32768 + 1024;
    B200_CHECK_CUDA(cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
    mma_bench_kernel<<<blocks, 128, 120 * 1024, stream>>>(ts_mode, n, iters, chains, out);
    (void)smem;
    B200_CHECK_CUDA(cudaGetLastError());
    return 0;
}
