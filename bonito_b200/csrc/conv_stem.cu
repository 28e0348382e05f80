// Fused conv stem of the LSTM-CRF encoder: Conv1d(1->C1,k=K1,'same') + act, then
// Conv1d(C1->C2,k=K2,'same') + act, written channels-last into a zero-padded buffer
//     out[n][PADL + l][c]   (l in [0,L), c in [0,C2)), Lp rows per chunk
// so that the strided conv that follows (k19 s6 in hac) is a plain GEMM over overlapping rows.
// The 164 MB (hac, batch 512) conv1 activation of the reference never reaches HBM.
// Reference semantics: bonito/nn.py:221-241 (Conv1d -> folded BN -> activation); the fp16
// rounding points of the reference's half-precision path are kept (conv output, then activation).
#include "common.cuh"

namespace {

constexpr int TL = 256;  // output positions per CTA

template <int C1, int K1, int C2, int K2>
__global__ void __launch_bounds__(TL)
conv_stem_kernel(const __half* __restrict__ x, int L, const __half* __restrict__ w1, const __half* __restrict__ b1,
                 int act1, const __half* __restrict__ w2, const __half* __restrict__ b2, int act2,
                 __half* __restrict__ out, int Lp, int padl) {
    constexpr int P1 = K1 / 2, P2 = K2 / 2;
    constexpr int NA1 = TL + K2 - 1;        // conv1 outputs needed by this tile
    constexpr int NX = NA1 + K1 - 1;        // input samples needed
    __shared__ float xs[NX];
    __shared__ float a1s[C1][NA1];
    __shared__ __align__(16) float w2s[C1 * K2][C2];
    __shared__ float w1s[K1][C1];
    __shared__ float b1s[C1], b2s[C2];

    const int tid = threadIdx.x;
    const int n = blockIdx.y;
    const int p0 = blockIdx.x * TL;   // first padded position of the tile
    const int l0 = p0 - padl;         // its signal coordinate
    const __half* xn = x + (long long)n * L;

    for (int i = tid; i < NX; i += TL) {
        int l = l0 - P2 - P1 + i;
        xs[i] = (l >= 0 && l < L) ? __half2float(xn[l]) : 0.f;
    }
    for (int i = tid; i < C1 * K2 * C2; i += TL) {
        int co = i % C2, ck = i / C2;  // ck = cin*K2 + tap
        int cin = ck / K2, tap = ck % K2;
        w2s[ck][co] = __half2float(w2[(co * C1 + cin) * K2 + tap]);
    }
    for (int i = tid; i < K1 * C1; i += TL) {
        int c = i % C1, k = i / C1;
        w1s[k][c] = __half2float(w1[c * K1 + k]);
    }
    if (tid < C1) b1s[tid] = b1 ? __half2float(b1[tid]) : 0.f;
    if (tid < C2) b2s[tid] = b2 ? __half2float(b2[tid]) : 0.f;
    __syncthreads();

    for (int i = tid; i < NA1; i += TL) {
        int l = l0 - P2 + i;
        bool in = (l >= 0 && l < L);
        float xv[K1];
#pragma unroll
        for (int k = 0; k < K1; ++k) xv[k] = xs[i + k];
#pragma unroll
        for (int c = 0; c < C1; ++c) {
            float acc = b1s[c];
#pragma unroll
            for (int k = 0; k < K1; ++k) acc = fmaf(w1s[k][c], xv[k], acc);
            a1s[c][i] = in ? apply_act_f16(acc, act1, 0.f, 0.f) : 0.f;
        }
    }
    __syncthreads();

    const int p = p0 + tid;
    if (p >= Lp) return;
    const int l = p - padl;
    __half* dst = out + ((long long)n * Lp + p) * C2;
    if (l < 0 || l >= L) {
#pragma unroll
        for (int c = 0; c < C2; c += 8) *reinterpret_cast<uint4*>(dst + c) = make_uint4(0, 0, 0, 0);
        return;
    }
    float acc[C2];
#pragma unroll
    for (int c = 0; c < C2; ++c) acc[c] = b2s[c];
#pragma unroll 4
    for (int cin = 0; cin < C1; ++cin) {
#pragma unroll
        for (int tap = 0; tap < K2; ++tap) {
            float v = a1s[cin][tid + tap];
            const float4* wrow = reinterpret_cast<const float4*>(w2s[cin * K2 + tap]);
#pragma unroll
            for (int q = 0; q < C2 / 4; ++q) {
                float4 w = wrow[q];
                acc[4 * q + 0] = fmaf(w.x, v, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(w.y, v, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(w.z, v, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(w.w, v, acc[4 * q + 3]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < C2; c += 8) {
        __half2 h[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            h[q] = __floats2half2_rn(apply_act_f16(acc[c + 2 * q], act2, 0.f, 0.f),
                                     apply_act_f16(acc[c + 2 * q + 1], act2, 0.f, 0.f));
        *reinterpret_cast<uint4*>(dst + c) = *reinterpret_cast<uint4*>(h);
    }
}

}  // namespace

int launch_conv_stem(const __half* x, int N, int L, int C1, int K1, const __half* w1, const __half* b1, int act1,
                     int C2, int K2, const __half* w2, const __half* b2, int act2, __half* out, int Lp, int padl,
                     cudaStream_t stream) {
    dim3 grid((Lp + TL - 1) / TL, N);
#define STEM_CASE(c1, k1, c2, k2)                                                                         \
    if (C1 == c1 && K1 == k1 && C2 == c2 && K2 == k2) {                                                   \
        conv_stem_kernel<c1, k1, c2, k2><<<grid, TL, 0, stream>>>(x, L, w1, b1, act1, w2, b2, act2, out,  \
                                                                  Lp, padl);                              \
        B200_CHECK_CUDA(cudaGetLastError());                                                              \
        return 0;                                                                                         \
    }
    STEM_CASE(16, 5, 16, 5)  // v4.x / v5.x fast, hac, sup LSTM models
    STEM_CASE(4, 5, 16, 5)   // old-style rnn_encoder (bonito/crf/model.py:150-162)
#undef STEM_CASE
    b200_set_error("conv_stem: unsupported shape 1->%d (k%d) -> %d (k%d)", C1, K1, C2, K2);
    return -2;
}
