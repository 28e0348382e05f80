// extern "C" surface of libbonito_b200.so (declared in include/bonito_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/bonito_b200.h"
#include "common.cuh"

int launch_conv_stem(const __half* x, int N, int L, int C1, int K1, const __half* w1, const __half* b1, int act1,
                     int C2, int K2, const __half* w2, const __half* b2, int act2, __half* out, int Lp, int padl,
                     cudaStream_t stream);
int lstm_rec_cluster_size(int H);
int launch_lstm_rec(const __half* gx, const __half* whh, __half* y, int T, int N, int H, int reverse,
                    cudaStream_t stream);
int launch_conv_first(const __half* x, int N, int L, int C, int K, const __half* w, const __half* bias, int act,
                      __half* out, int Lp, int padl, cudaStream_t stream);
int launch_rmsnorm_residual(const __half* a, const __half* x, const __half* w, float alpha, float eps, __half* out,
                            long long M, int D, cudaStream_t stream);
int launch_swiglu(const __half* h, __half* out, long long M, int F, cudaStream_t stream);
int launch_attention(__half* qkv, const __half* cos_sin, __half* out, int N, int T, int NH, int head_dim, int wl,
                     int wr, cudaStream_t stream);
bool lstm_rec_tc_supported(int hidden);
int launch_lstm_rec_tc(const __half* gx, const __half* whh, __half* y, int T, int N, int hidden, int reverse,
                       cudaStream_t stream);
int launch_tmem_probe(float* out, cudaStream_t stream);
int lstm_rec_tile_chunks(int hidden);
int lstm_rec_tile_cluster(int hidden);
int launch_lstm_rec_tc6(const __half* gx, const __half* whh, __half* y, void* workspace, int T, int N, int hidden,
                        int reverse, cudaStream_t stream);
size_t lstm_rec_tile_workspace_bytes(int N);
int copy_lstm_timeline6(long long* host_out, int max_steps);
int copy_lstm_timeline(long long* host_out, int max_steps);
int lstm_rec_tc_max_clusters();
int debug_max_clusters(int cluster_size, int threads, int smem_bytes);
int launch_exchange_bench(int mode, int steps, int delay, int clusters, unsigned char* staging, long long* out,
                          cudaStream_t stream);
int launch_mma_bench(int ts_mode, int n, int iters, int chains, int blocks, long long* out, cudaStream_t stream);
size_t crf_decode_workspace_bytes(int N, int T, int state_len);
int launch_crf_decode(const __half* scores, int N, int T, int state_len, float blank, float qscale, float qbias,
                      void* workspace, uint8_t* moves, uint8_t* seq, uint8_t* qual, cudaStream_t stream);

// one message buffer per host thread: the reference drives this path from background threads (bonito/multiprocessing.py:118-122),
// so a failing call must read back its own message, not another thread's
int launch_crf_beam_search(const __half* scores, int N, int T, int state_len, float blank, int width, float cut, float qscale,
                           float qbias, void* workspace, uint8_t* moves, uint8_t* seq, uint8_t* qual, cudaStream_t stream);

int launch_lstm_crf_fwd(const b200_lstm_crf_plan* p, const __half* x, __half* scores, cudaStream_t stream);

int copy_attention_timeline(long long* host_out, int max_tiles);

int launch_quantize_i8(const __half* x, int8_t* out, long long n, float scale, cudaStream_t stream);
int launch_gemm_i8(const int8_t* A, long long lda, const int8_t* B, const float* col_scale, __half* C, long long ldc, int M,
                   int N, int K, const GemmEpilogue& ep, int max_ctas, cudaStream_t stream);

static thread_local char g_err[1024] = "";

void b200_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {

int b200_version(void) { return 100; }

const char* b200_last_error(void) { return g_err; }

int b200_conv_stem_fwd(const void* x, int n, int l, int c1, int k1, const void* w1, const void* b1, int act1,
                       int c2, int k2, const void* w2, const void* b2, int act2, void* out, int lp, int padl,
                       void* stream) {
    B200_REQUIRE(x && w1 && w2 && out, "conv_stem: null pointer argument");
    B200_REQUIRE(n >= 0 && l > 0 && lp >= padl + l, "conv_stem: bad sizes n=%d l=%d lp=%d padl=%d", n, l, lp, padl);
    if (n == 0) return 0;
    return launch_conv_stem((const __half*)x, n, l, c1, k1, (const __half*)w1, (const __half*)b1, act1, c2, k2,
                            (const __half*)w2, (const __half*)b2, act2, (__half*)out, lp, padl,
                            (cudaStream_t)stream);
}

int b200_gemm_fwd(const void* a, long long lda, const void* b, const void* bias, void* c, long long ldc, int m,
                  int n, int k, int act, float lo, float hi, int rows_inner, int valid_inner,
                  long long stride_inner, long long stride_outer, int impl, void* stream) {
    return b200_gemm_fwd_ex(a, lda, b, bias, c, ldc, m, n, k, act, lo, hi, rows_inner, valid_inner, stride_inner,
                            stride_outer, 0, 0, 0, 0, impl, 0, stream);
}

int b200_gemm_fwd_ex(const void* a, long long lda, const void* b, const void* bias, void* c, long long ldc, int m,
                     int n, int k, int act, float lo, float hi, int rows_inner, int valid_inner,
                     long long stride_inner, long long stride_outer, int group, long long stride_group, int cb_width,
                     int cb_rows, int impl, int max_ctas, void* stream) {
    B200_REQUIRE(a && b && c, "gemm: null pointer argument");
    B200_REQUIRE(m >= 0 && n > 0 && k > 0 && rows_inner > 0 && group >= 0, "gemm: bad sizes m=%d n=%d k=%d", m, n, k);
    B200_REQUIRE(k % 8 == 0 && lda % 8 == 0 && n % 8 == 0 && ldc % 8 == 0,
                 "gemm: k, lda, n, ldc must be multiples of 8 (k=%d lda=%lld n=%d ldc=%lld)", k, lda, n, ldc);
    B200_REQUIRE(cb_width >= 0 && cb_rows >= 0 && cb_width % 32 == 0 && (cb_width == 0 || act != B200_ACT_SWIGLU),
                 "gemm: column blocks must be multiples of 32 columns and cannot be combined with SwiGLU (cb_width=%d)",
                 cb_width);
    if (act == B200_ACT_SWIGLU)
        B200_REQUIRE(n % 64 == 0 && !bias && impl != B200_GEMM_MMA_SYNC,
                     "gemm: the fused SwiGLU epilogue needs n %% 64 == 0, no bias and the tcgen05 path (n=%d)", n);
    if (m == 0) return 0;
    GemmEpilogue ep;
    ep.bias = (const __half*)bias;
    ep.act = act;
    ep.lo = lo;
    ep.hi = hi;
    ep.map.rows_inner = rows_inner;
    ep.map.valid_inner = valid_inner;
    ep.map.stride_inner = stride_inner;
    ep.map.stride_outer = stride_outer;
    ep.map.group = group;
    ep.map.stride_group = stride_group;
    ep.cb_width = cb_width;
    ep.cb_rows = cb_rows;
    if (impl == B200_GEMM_AUTO) {
        const char* env = getenv("B200_GEMM_IMPL");
        impl = (env && strcmp(env, "mma") == 0 && act != B200_ACT_SWIGLU) ? B200_GEMM_MMA_SYNC : B200_GEMM_TCGEN05;
    }
    if (impl == B200_GEMM_MMA_SYNC)
        return launch_gemm_mma((const __half*)a, lda, (const __half*)b, (__half*)c, ldc, m, n, k, ep,
                               (cudaStream_t)stream);
    return launch_gemm_tc((const __half*)a, lda, (const __half*)b, (__half*)c, ldc, m, n, k, ep, max_ctas,
                          (cudaStream_t)stream, impl == B200_GEMM_TCGEN05_PAIR);
}

int b200_conv_first_fwd(const void* x, int n, int l, int c, int k, const void* w, const void* bias, int act, void* out,
                        int lp, int padl, void* stream) {
    B200_REQUIRE(x && w && out, "conv_first: null pointer argument");
    B200_REQUIRE(n >= 0 && l > 0 && lp >= padl + l, "conv_first: bad sizes n=%d l=%d lp=%d padl=%d", n, l, lp, padl);
    if (n == 0) return 0;
    return launch_conv_first((const __half*)x, n, l, c, k, (const __half*)w, (const __half*)bias, act, (__half*)out, lp,
                             padl, (cudaStream_t)stream);
}

int b200_attention_fwd(void* qkv, const void* cos_sin, void* out, int n, int t, int heads, int head_dim, int wl,
                       int wr, void* stream) {
    B200_REQUIRE(qkv && cos_sin && out, "attention: null pointer argument");
    B200_REQUIRE(n >= 0 && t >= 0 && heads > 0, "attention: bad sizes n=%d t=%d heads=%d", n, t, heads);
    if (n == 0 || t == 0) return 0;
    return launch_attention((__half*)qkv, (const __half*)cos_sin, (__half*)out, n, t, heads, head_dim, wl, wr,
                            (cudaStream_t)stream);
}

int b200_rmsnorm_residual_fwd(const void* a, const void* x, const void* w, float alpha, float eps, void* out,
                              long long m, int d, void* stream) {
    B200_REQUIRE(a && x && w && out, "rmsnorm: null pointer argument");
    if (m <= 0) return 0;
    return launch_rmsnorm_residual((const __half*)a, (const __half*)x, (const __half*)w, alpha, eps, (__half*)out, m, d,
                                   (cudaStream_t)stream);
}

int b200_swiglu_fwd(const void* h, void* out, long long m, int f, void* stream) {
    B200_REQUIRE(h && out, "swiglu: null pointer argument");
    if (m <= 0) return 0;
    return launch_swiglu((const __half*)h, (__half*)out, m, f, (cudaStream_t)stream);
}

int b200_lstm_cluster_size(int hidden) { return lstm_rec_cluster_size(hidden); }

int b200_lstm_rec_fwd(const void* gx, const void* whh, void* y, int t, int n, int hidden, int reverse,
                      void* stream) {
    B200_REQUIRE(gx && whh && y, "lstm_rec: null pointer argument");
    B200_REQUIRE(t >= 0 && n >= 0, "lstm_rec: bad sizes t=%d n=%d", t, n);
    if (t == 0 || n == 0) return 0;
    const char* env = getenv("B200_LSTM_IMPL");
    const bool force_mma = env && strcmp(env, "mma") == 0;
    if (!force_mma && lstm_rec_tc_supported(hidden))
        return launch_lstm_rec_tc((const __half*)gx, (const __half*)whh, (__half*)y, t, n, hidden, reverse,
                                  (cudaStream_t)stream);
    return launch_lstm_rec((const __half*)gx, (const __half*)whh, (__half*)y, t, n, hidden, reverse,
                           (cudaStream_t)stream);
}

int b200_lstm_tile_chunks(int hidden) { return lstm_rec_tile_chunks(hidden); }

int b200_lstm_tile_cluster(int hidden) { return lstm_rec_tile_cluster(hidden); }

size_t b200_lstm_rec_tile_workspace_bytes(int n) { return lstm_rec_tile_workspace_bytes(n); }

int b200_lstm_rec_tile_fwd(const void* gx, const void* whh, void* y, void* workspace, int t, int n, int hidden,
                           int reverse, void* stream) {
    B200_REQUIRE(gx && whh && y && workspace, "lstm_rec_tile: null pointer argument");
    B200_REQUIRE(t >= 0 && n >= 0, "lstm_rec_tile: bad sizes t=%d n=%d", t, n);
    if (t == 0 || n == 0) return 0;
    return launch_lstm_rec_tc6((const __half*)gx, (const __half*)whh, (__half*)y, workspace, t, n, hidden, reverse,
                               (cudaStream_t)stream);
}

int b200_debug_lstm_tile_timeline(long long* host_out, int max_steps) {
    B200_REQUIRE(host_out != nullptr && max_steps > 0, "lstm_tile_timeline: bad arguments");
    return copy_lstm_timeline6(host_out, max_steps);
}

int b200_debug_attention_timeline(long long* host_out, int max_tiles) {
    B200_REQUIRE(host_out != nullptr && max_tiles > 0, "attention_timeline: bad arguments");
    return copy_attention_timeline(host_out, max_tiles);
}

int b200_debug_gemm_profile(long long* host_out) {
    B200_REQUIRE(host_out != nullptr, "gemm_profile: null pointer argument");
    return copy_gemm_profile(host_out);
}

int b200_debug_tmem_probe(void* out, void* stream) {
    B200_REQUIRE(out != nullptr, "tmem_probe: null pointer argument");
    return launch_tmem_probe((float*)out, (cudaStream_t)stream);
}

int b200_debug_lstm_max_clusters(void) { return lstm_rec_tc_max_clusters(); }

int b200_debug_exchange_bench(int mode, int steps, int delay, int clusters, void* staging, void* out, void* stream) {
    B200_REQUIRE(staging && out && steps > 0 && clusters > 0 && clusters <= 22 && (mode == 0 || (mode >= 2 && mode <= 4)),
                 "exchange_bench: bad arguments");
    return launch_exchange_bench(mode, steps, delay, clusters, (unsigned char*)staging, (long long*)out, (cudaStream_t)stream);
}

int b200_debug_max_clusters(int cluster_size, int threads, int smem_bytes) {
    return debug_max_clusters(cluster_size, threads, smem_bytes);
}

int b200_debug_lstm_timeline(long long* host_out, int max_steps) {
    B200_REQUIRE(host_out != nullptr && max_steps > 0, "lstm_timeline: bad arguments");
    return copy_lstm_timeline(host_out, max_steps);
}

int b200_debug_mma_bench(int ts_mode, int n, int iters, int chains, int blocks, void* out, void* stream) {
    B200_REQUIRE(out != nullptr && n >= 16 && n <= 256 && n % 16 == 0 && iters > 0 && chains >= 1 && chains * n <= 448,
                 "mma_bench: bad arguments");
    return launch_mma_bench(ts_mode, n, iters, chains, blocks, (long long*)out, (cudaStream_t)stream);
}

size_t b200_crf_decode_workspace_bytes(int n, int t, int state_len) {
    return crf_decode_workspace_bytes(n, t, state_len);
}

int b200_crf_decode(const void* scores, int n, int t, int state_len, float blank_score, float qscale, float qbias,
                    void* workspace, void* moves, void* sequence, void* qstring, void* stream) {
    B200_REQUIRE(n >= 0 && t >= 0, "crf_decode: bad sizes n=%d t=%d", n, t);
    if (n == 0 || t == 0) return 0;
    B200_REQUIRE(scores && workspace && moves && sequence && qstring, "crf_decode: null pointer argument");
    return launch_crf_decode((const __half*)scores, n, t, state_len, blank_score, qscale, qbias, workspace,
                             (uint8_t*)moves, (uint8_t*)sequence, (uint8_t*)qstring, (cudaStream_t)stream);
}

int b200_chunk_count(long long length, int chunksize, int overlap) {
    if (length <= 0 || chunksize <= 0 || overlap < 0 || overlap >= chunksize) return 0;
    return chunk_count(length, chunksize, overlap);
}

int b200_chunk_signal(const void* signal, int signal_is_f32, long long length, int chunksize, int overlap, void* out,
                      long long row_stride, void* stream) {
    B200_REQUIRE(signal && out, "chunk_signal: null pointer argument");
    return launch_chunk_signal(signal, signal_is_f32, length, chunksize, overlap, (__half*)out, row_stride, (cudaStream_t)stream);
}

int b200_stream_create(void** stream_out) {
    B200_REQUIRE(stream_out != nullptr, "stream_create: null pointer argument");
    cudaStream_t st = nullptr;
    B200_CHECK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *stream_out = (void*)st;
    return 0;
}

int b200_quantize_i8(const void* x, void* out, long long n, float scale, void* stream) {
    B200_REQUIRE(x && out && n >= 0, "quantize_i8: bad arguments");
    return launch_quantize_i8((const __half*)x, (int8_t*)out, n, scale, (cudaStream_t)stream);
}

int b200_gemm_i8_fwd(const void* a, long long lda, const void* b, const void* col_scale, const void* bias, void* c,
                     long long ldc, int m, int n, int k, int act, float lo, float hi, int rows_inner, int valid_inner,
                     long long stride_inner, long long stride_outer, int group, long long stride_group, int cb_width,
                     int cb_rows, int max_ctas, void* stream) {
    B200_REQUIRE(a && b && c && col_scale, "gemm_i8: null pointer argument");
    B200_REQUIRE(m >= 0 && n > 0 && k > 0 && rows_inner > 0 && group >= 0 && n % 8 == 0 && ldc % 8 == 0,
                 "gemm_i8: bad sizes m=%d n=%d k=%d", m, n, k);
    B200_REQUIRE(cb_width >= 0 && cb_rows >= 0 && cb_width % 32 == 0, "gemm_i8: column blocks must be multiples of 32 columns");
    if (m == 0) return 0;
    GemmEpilogue ep;
    ep.bias = (const __half*)bias;
    ep.act = act;
    ep.lo = lo;
    ep.hi = hi;
    ep.map = RowMap{rows_inner, valid_inner, stride_inner, stride_outer, group, stride_group};
    ep.cb_width = cb_width;
    ep.cb_rows = cb_rows;
    return launch_gemm_i8((const int8_t*)a, lda, (const int8_t*)b, (const float*)col_scale, (__half*)c, ldc, m, n, k, ep, max_ctas,
                          (cudaStream_t)stream);
}

int b200_lstm_crf_fwd(const b200_lstm_crf_plan* plan, const void* x, void* scores, void* stream) {
    return launch_lstm_crf_fwd(plan, (const __half*)x, (__half*)scores, (cudaStream_t)stream);
}

int b200_crf_beam_search(const void* scores, int n, int t, int state_len, float blank_score, int beam_width, float beam_cut,
                         float qscale, float qbias, void* workspace, void* moves, void* sequence, void* qstring, void* stream) {
    B200_REQUIRE(n >= 0 && t >= 0, "beam_search: bad sizes n=%d t=%d", n, t);
    if (n == 0 || t == 0) return 0;
    B200_REQUIRE(scores && workspace && moves && sequence && qstring, "beam_search: null pointer argument");
    return launch_crf_beam_search((const __half*)scores, n, t, state_len, blank_score, beam_width, beam_cut, qscale, qbias,
                                  workspace, (uint8_t*)moves, (uint8_t*)sequence, (uint8_t*)qstring, (cudaStream_t)stream);
}

}  // extern "C"
