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

// dummy cluster kernel for occupancy queries
__global__ void cluster_probe_kernel(int* out) {
    extern __shared__ unsigned char smem_raw[];
    if (threadIdx.x == 0 && out != nullptr) atomicAdd(out, (int)cluster_ctarank() + (smem_raw[0] & 0));
}

}  // namespace

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
