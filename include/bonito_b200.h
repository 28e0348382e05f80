/*
 * bonito_b200 -- C ABI of the B200-native chunked forward + decode path.
 *
 * The reference (nanoporetech/bonito) is pure Python; its native work on this path is done by
 * third-party binaries reached from these call sites, which are what each entry point replaces:
 *
 *   b200_conv_stem_fwd        torch.nn.Conv1d x2 + activations        bonito/nn.py:221-241
 *   b200_gemm_fwd             torch.nn.Conv1d (strided, as GEMM),     bonito/nn.py:226,283-298,59-67
 *                             torch.nn.Linear + Clamp (LinearCRFEncoder), LSTM input projection
 *   b200_lstm_rec_fwd         koi.lstm.update_graph / torch.nn.LSTM   bonito/crf/model.py:240-246, bonito/nn.py:366-370
 *   b200_crf_decode           koi.decode.beam_search call contract    bonito/crf/basecall.py:36-40
 *                             with SeqdistModel.decode_batch maths    bonito/crf/model.py:98-108,196-199
 *
 * Conventions: every function returns 0 on success and a negative value on
 * failure, with a message available from b200_last_error().  All pointers are raw DEVICE pointers
 * owned by the caller (fp16 = IEEE binary16); the library never allocates or frees caller memory and
 * keeps no thread-local CUDA state: work is enqueued on the `stream` argument (a cudaStream_t passed
 * as void*) of the device that is current on the calling thread.  Workspace sizes come from the
 * *_workspace_bytes() queries.  Safe to call from any host thread.
 */
#ifndef BONITO_B200_H
#define BONITO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ACT_NONE 0
#define B200_ACT_SWISH 1
#define B200_ACT_TANH 2
#define B200_ACT_CLAMP 3 /* clamp(lo, hi), Clamp layer: bonito/nn.py:59-67 */
#define B200_ACT_SCALE 4 /* multiply by lo, LinearCRFEncoder.scale: bonito/nn.py:288-289 */
/* SwiGLU fused into the GEMM (tcgen05 path only): the n output columns are 64-wide groups [32 x y | 32 x gate] (the caller
 * interleaves the rows of fc1.weight that way) and c receives n/2 columns, c[:, 32*g + j] = gate * y / (1 + exp(-gate)) on the
 * fp16-rounded y / gate, rounded once -- GatedMlp: flash_attn/modules/mlp.py:99-136, flash_attn/ops/activations.py:107-111
 * as used by bonito/transformer/model.py:100-104.  n % 64 == 0, no bias. */
#define B200_ACT_SWIGLU 5
#define B200_ACT_TANH_SCALE 6 /* tanh, then multiply by lo: LinearCRFEncoder(activation="tanh", scale=5.0), bonito/nn.py:283-298 */

#define B200_GEMM_AUTO 0 /* tcgen05 (product path) unless B200_GEMM_IMPL=mma is set in the environment */
#define B200_GEMM_TCGEN05 1
#define B200_GEMM_MMA_SYNC 2 /* legacy tensor path, kept for on-device cross-checks */
#define B200_GEMM_TCGEN05_PAIR 3 /* cta_group::2 pair kernels (N % 256 == 0): faster alone, not the default next to other kernels */

/* Library version (major*10000 + minor*100 + patch). */
int b200_version(void);

/* Message of the last failure ON THE CALLING THREAD ("" if none); the buffer is thread-local and stays valid until the
 * same thread fails again. */
const char* b200_last_error(void);

/*
 * Fused conv stem: x[N][L] -> Conv1d(1->c1,k1,pad k1/2)+act1 -> Conv1d(c1->c2,k2,pad k2/2)+act2,
 * written channels-last and zero padded: out[n][padl + l][c], `lp` rows per chunk (rows outside
 * [padl, padl+L) are written as zeros).  Weights in torch layout (w1 [c1][1][k1], w2 [c2][c1][k2]),
 * biases may be NULL.  Supported shapes: (c1,k1,c2,k2) = (16,5,16,5), (4,5,16,5).
 */
int b200_conv_stem_fwd(const void* x, int n, int l, int c1, int k1, const void* w1, const void* b1, int act1,
                       int c2, int k2, const void* w2, const void* b2, int act2, void* out, int lp, int padl,
                       void* stream);

/*
 * C = act(A[M,K] * B[N,K]^T + bias) in fp16 with fp32 accumulation.
 *   A: row stride `lda` elements (lda < K is allowed: overlapping rows = strided convolution windows)
 *   B: [N][K] row-major (torch Linear / packed Conv1d weight), bias[N] or NULL
 *   output row of input row r: (outer, inner) = divmod(r, rows_inner); rows with inner >= valid_inner are
 *   skipped; out_row = inner*stride_inner + outer*stride_outer; C row stride `ldc` elements.
 *   (identity mapping: rows_inner = valid_inner = M, stride_inner = 1, stride_outer = 0)
 * Requirements: K % 8 == 0, lda % 8 == 0, N % 8 == 0, ldc % 8 == 0, A/B/C 16-byte aligned.
 */
int b200_gemm_fwd(const void* a, long long lda, const void* b, const void* bias, void* c, long long ldc, int m,
                  int n, int k, int act, float lo, float hi, int rows_inner, int valid_inner,
                  long long stride_inner, long long stride_outer, int impl, void* stream);

/*
 * Same, with
 *   max_ctas  a cap on the number of (persistent) CTAs the tcgen05 kernel may occupy (0 = every SM): the tile-pipelined
 *             engine caps the GEMMs it runs next to resident recurrent clusters;
 *   group, stride_group  second level of the row map (group = 0: off): (outer2, outer1) = divmod(outer, group),
 *             out_row = inner*stride_inner + outer1*stride_outer + outer2*stride_group -- e.g. chunk n -> (tile n/48, n%48);
 *   cb_width, cb_rows  column blocks (cb_width = 0: off; multiple of 32): output column c of mapped row R is written to
 *             row R + (c / cb_width) * cb_rows, column c % cb_width.  With ldc = cb_width this lays the LSTM input
 *             projection out as [t][cluster rank][chunk][cb_width], one contiguous block per recurrent CTA and step.
 */
int b200_gemm_fwd_ex(const void* a, long long lda, const void* b, const void* bias, void* c, long long ldc, int m,
                     int n, int k, int act, float lo, float hi, int rows_inner, int valid_inner,
                     long long stride_inner, long long stride_outer, int group, long long stride_group, int cb_width,
                     int cb_rows, int impl, int max_ctas, void* stream);

/*
 * Cluster size the packed LSTM operands must be laid out for (0: hidden size unsupported).
 * Supported hidden sizes: 96, 128, 256, 384.
 */
int b200_lstm_cluster_size(int hidden);

/*
 * Recurrent part of one unidirectional LSTM layer over all T steps (h0 = c0 = 0):
 *   gx  [T][N][4H]  input projection x_t W_ih^T + b_ih + b_hh, columns permuted to
 *                   [cluster rank][unit/8 block][unit%8][gate i,f,g,o]
 *   whh [4H][H]     recurrent weights, rows permuted to [cluster rank][unit/8 block][gate][unit%8]
 *   y   [T][N][H]   h_t in natural unit order
 * reverse != 0 runs t = T-1..0 (the reference flips the sequence instead: bonito/nn.py:366-370).
 * hidden = 384 runs the tcgen05 kernel (W_hh resident in tensor memory, h exchanged through distributed shared
 * memory); other sizes, or B200_LSTM_IMPL=mma in the environment, run the mma.sync kernel.
 */
int b200_lstm_rec_fwd(const void* gx, const void* whh, void* y, int t, int n, int hidden, int reverse,
                      void* stream);

/*
 * Tile layout of the H = 384 recurrent kernel (second generation: clusters of 6 CTAs x 64 hidden units, three interleaved
 * 16-chunk sub-tiles, input projection streamed through shared memory by cp.async.bulk).
 *   b200_lstm_tile_chunks(hidden)   chunks per tile (48 for hidden = 384; 0 = this hidden size has no tile kernel)
 *   b200_lstm_tile_cluster(hidden)  CTAs per cluster = column blocks of gx (6)
 *   gx  [tiles][T][6][48][256]  columns of cluster rank r: [unit/8 - 8r][unit%8][gate i,f,g,o]  (b200_gemm_fwd_ex with
 *                               rows (t, chunk), cb_width = 256, cb_rows = 48, ldc = 256)
 *   whh [4H][H]                 as for b200_lstm_rec_fwd
 *   y   [tiles][T][48][H]       h_t in natural unit order; rows of chunks >= n are not written
 *   workspace                   b200_lstm_rec_tile_workspace_bytes(n) bytes (72 KB per tile): staging of the h all-gather,
 *                               which goes through L2 as multicast bulk copies; contents irrelevant, but launches that
 *                               may run concurrently need distinct workspaces
 * tiles = ceil(n / 48); one launch runs all of them (one cluster each; 22 fit on a B200 at once).
 */
int b200_lstm_tile_chunks(int hidden);
int b200_lstm_tile_cluster(int hidden);
size_t b200_lstm_rec_tile_workspace_bytes(int n);
int b200_lstm_rec_tile_fwd(const void* gx, const void* whh, void* y, void* workspace, int t, int n, int hidden,
                           int reverse, void* stream);

/* Timing aid: after a b200_attention_fwd launched with B200_ATTN_DEBUG=1, the SM-clock stamps CTA 0 recorded for its first
 * query tiles ([tile][16] int64, HOST buffer; see attention_tc.cu).  Returns the number of tiles copied (<= 64). */
int b200_debug_attention_timeline(long long* host_out, int max_tiles);
/* B200_GEMM_DEBUG=1: per-CTA cycle counters of the last weight-stationary GEMM launch, 160 x 8 values (gemm_tc.cu) */
int b200_debug_gemm_profile(long long* host_out);

/* Timing aid: as b200_debug_lstm_timeline, for b200_lstm_rec_tile_fwd. */
int b200_debug_lstm_tile_timeline(long long* host_out, int max_steps);

/*
 * ---- transformer (sup) path: bonito/transformer/model.py ----
 *
 * First convolution of a conv stack: x[N][L] -> Conv1d(1->c, k, pad k/2) + act, channels-last with zero halo:
 * out[n][padl + l][c], `lp` rows per chunk.  The following convolutions run as b200_gemm_fwd over overlapping rows.
 */
int b200_conv_first_fwd(const void* x, int n, int l, int c, int k, const void* w, const void* bias, int act, void* out,
                        int lp, int padl, void* stream);

/*
 * Rotary embedding (NeoX half rotation, cos_sin [T][64] fp16 = cos[32] | sin[32] per position) + windowed softmax
 * attention, non-causal: key j is visible to query i iff i - wl <= j <= i + wr (negative = unlimited).
 * qkv [N][T][3][heads][64] fp16 (packed projection, bonito/transformer/model.py:71) -> out [N][T][heads*64].
 * The q and k parts of `qkv` are rotated IN PLACE (as flash-attn's RotaryEmbedding does, transformer/model.py:73).
 */
int b200_attention_fwd(void* qkv, const void* cos_sin, void* out, int n, int t, int heads, int head_dim, int wl,
                       int wr, void* stream);

/* out[r] = rmsnorm(a[r] + fp16(alpha * x[r]), eps) * w   for m rows of d elements (DeepNorm post-norm residual). */
int b200_rmsnorm_residual_fwd(const void* a, const void* x, const void* w, float alpha, float eps, void* out,
                              long long m, int d, void* stream);

/* h [m][2f] = (y | gate) -> out [m][f] = gate * y / (1 + exp(-gate))   (GatedMlp with SiLU). */
int b200_swiglu_fwd(const void* h, void* out, long long m, int f, void* stream);

/*
 * Self-test of the tensor-memory conventions the tcgen05 kernels rely on (fragment layout of
 * tcgen05.ld.16x256b, fp16-pair packing of a TMEM-resident A operand, un-swizzled B tiles).  out: 16384 floats (device);
 * interpreted by tests/test_gpu_kernels.py::test_tmem_conventions.
 */
int b200_debug_tmem_probe(void* out, void* stream);

/* Number of 8-CTA clusters of the tcgen05 recurrent kernel the current device can hold at once (-1 on error). */
int b200_debug_lstm_max_clusters(void);

/*
 * Timing aid: the h all-gather of the tile recurrent kernel without the math (6-CTA clusters, 3 x 8 sender warps per CTA,
 * 256-byte blocks into the h tiles of all six CTAs every step; see debug_bench.cu for the modes: 0 DSMEM bulk copies,
 * 2 / 3 multicast bulk copies out of an L2 staging buffer, 4 DSMEM with 2 KB copies).  staging: clusters * 73728 bytes
 * (device); out: 2 x int64 (device) = cycles of CTA 0, steps.
 */
int b200_debug_exchange_bench(int mode, int steps, int delay, int clusters, void* staging, void* out, void* stream);

/* Occupancy query: clusters of `cluster_size` CTAs (`threads` threads, `smem_bytes` dynamic shared memory, one CTA per SM
 * when smem_bytes > half an SM) the current device holds at once; -1 on error.  GPC packing decides (B200: 148 SMs). */
int b200_debug_max_clusters(int cluster_size, int threads, int smem_bytes);

/*
 * Timing aid: after a b200_lstm_rec_fwd launched with B200_LSTM_DEBUG=3 in the environment, copies the SM-clock
 * stamps CTA 0 recorded for its first steps ([step][8] int64, HOST buffer; see lstm_rec_tc.cu).  Returns the
 * number of steps copied (<= 256) or a negative error.
 */
int b200_debug_lstm_timeline(long long* host_out, int max_steps);

/*
 * Timing aid: `iters` tcgen05.mma (M=128, N=n, K=16, fp16) round-robin over `chains` independent accumulators, A from
 * tensor memory (ts_mode=1) or shared memory (0), on `blocks` CTAs; out (device, 3 x int64): issue cycles,
 * issue-to-completion cycles, nanoseconds.
 */
int b200_debug_mma_bench(int ts_mode, int n, int iters, int chains, int blocks, void* out, void* stream);

/* Bytes of scratch b200_crf_decode needs for n chunks of t frames. */
size_t b200_crf_decode_workspace_bytes(int n, int t, int state_len);

/*
 * Posterior + Viterbi decode of CRF scores.
 *   scores [N][T][4^(state_len+1)] fp16, no blank column (index = state*4 + dropped_base);
 *   blank_score: fixed stay score (reference LinearCRFEncoder.blank_score / beam_search default 2.0)
 *   moves/sequence/qstring: [N][T] bytes; sequence/qstring hold an ASCII char on move frames and 0
 *   elsewhere, so `to_str` = bytes of the non-zero entries (bonito/crf/basecall.py:50-54);
 *   quality = phred of the posterior move mass of the emitted base, q = -10 log10(max(1-p,1e-4))*qscale+qbias.
 */
int b200_crf_decode(const void* scores, int n, int t, int state_len, float blank_score, float qscale, float qbias,
                    void* workspace, void* moves, void* sequence, void* qstring, void* stream);

/*
 * ---- chunk() on the device (reference: bonito.util.chunk, bonito/util.py:142-161) ----
 * b200_chunk_count: number of chunks of a read of `length` samples (0 for an invalid geometry).
 * b200_chunk_signal: signal [length] (fp16, or fp32 when signal_is_f32) -> out [b200_chunk_count][chunksize] fp16, rows
 *   `row_stride` elements apart: reads shorter than a chunk are tiled, otherwise windows every chunksize - overlap samples from
 *   stub = (length - overlap) % (chunksize - overlap), preceded by signal[:chunksize] when stub > 0.
 */
int b200_chunk_count(long long length, int chunksize, int overlap);
int b200_chunk_signal(const void* signal, int signal_is_f32, long long length, int chunksize, int overlap, void* out,
                      long long row_stride, void* stream);

/*
 * b200_stream_create: a non-blocking CUDA stream on the current device, for the lifetime of the process (host frameworks
 * that hand out pooled streams -- torch: 32 per device, round-robin -- cannot promise that two streams are distinct; the
 * pipelined host loop needs its per-batch and copy streams to be).
 */
int b200_stream_create(void** stream_out);

/*
 * ---- INT8 input projection (--quantize; reference: koi's int8 LSTM path, bonito/crf/model.py:245, cli/basecaller.py:186-189) ----
 * b200_quantize_i8: out[i] = clamp(rint(x[i] * scale), -127, 127), fp16 -> int8, n a multiple of 8.
 * b200_gemm_i8_fwd: C = act(col_scale[j] * sum_k A_i8[i][k] B_i8[j][k] + bias[j]) -- int8 operands (lda in bytes), s32
 *   accumulation on tcgen05 kind::i8, per-column float scale (weight scale / activation scale), fp16 bias / output, the
 *   same row / column-block maps as b200_gemm_fwd_ex.  K <= 768, K % 16 == 0, N a multiple of 192 or 128.
 */
int b200_quantize_i8(const void* x, void* out, long long n, float scale, void* stream);
int b200_gemm_i8_fwd(const void* a, long long lda, const void* b, const void* col_scale, const void* bias, void* c,
                     long long ldc, int m, int n, int k, int act, float lo, float hi, int rows_inner, int valid_inner,
                     long long stride_inner, long long stride_outer, int group, long long stride_group, int cb_width,
                     int cb_rows, int max_ctas, void* stream);

/*
 * ---- coarse entry point: the whole LSTM-CRF encoder forward of one batch from one call ----
 * conv stem -> strided convolution (GEMM) -> n_lstm x (input projection GEMM + persistent recurrent layer) ->
 * LinearCRFEncoder GEMM (+Clamp), enqueued on `stream` (14 launches for the hac shape).  Replaces the module-tree walk of
 * `Serial.forward` over the encoder of a bonito.crf model (bonito/nn.py:82-89, bonito/crf/model.py:150-162) -- the span
 * `Model.use_koi` swaps for koi.lstm.update_graph plus the layers around it.  Tile-layout recurrent kernel only
 * (b200_lstm_tile_chunks(hidden) > 0).  The plan holds DEVICE pointers to packed weights (layouts as documented for the
 * fine-grained entry points above: conv weights in torch layout, w3 [H][k3*c2] with k = tap*c2 + cin, wih / bias in gx column
 * order, whh in W_hh row order) and to caller-owned work buffers:
 *   stem  (n*tp*s3*c2 + k3*c2) halves, the last k3*c2 zeroed       ya, yb  tiles*t*48*H halves, zero-filled once
 *   gx    tiles*t*4H*48 halves, zero-filled once                    hx      b200_lstm_rec_tile_workspace_bytes(n) bytes
 * with tiles = ceil(n / 48), t = frames, tp = padded frames per chunk (the stem buffer holds tp*s3 samples per chunk).
 * x [n][l] fp16 -> scores [n][t][n_scores] fp16 (no blank column).
 */
#define B200_MAX_LSTM_LAYERS 8
typedef struct b200_lstm_crf_plan {
    int n, l, t, tp;
    int c1, k1, act1, c2, k2, act2;          /* conv stem */
    int hidden, k3, s3, pad3, act3;          /* strided convolution into the LSTM width */
    int n_lstm, n_scores, act_l;
    float lo, hi;                            /* clamp bounds (act_l = B200_ACT_CLAMP) */
    int reverse[B200_MAX_LSTM_LAYERS];
    const void *w1, *b1, *w2, *b2, *w3, *b3, *wl, *bl;
    const void* wih[B200_MAX_LSTM_LAYERS];
    const void* bias[B200_MAX_LSTM_LAYERS];
    const void* whh[B200_MAX_LSTM_LAYERS];
    void *stem, *ya, *yb, *gx, *hx;
} b200_lstm_crf_plan;
int b200_lstm_crf_fwd(const b200_lstm_crf_plan* plan, const void* x, void* scores, void* stream);

/*
 * Beam-search decode with the argument meaning of koi.decode.beam_search (bonito/crf/basecall.py:36-40): beam_width entries
 * (1..32), candidates more than beam_cut (natural-log units) below the best are dropped.  koi itself is a closed binary with
 * no pinned outputs, so this is this library's own backward-guided prefix beam search (oracle: crf_oracle.beam_search_native);
 * the exact posterior-Viterbi decoder b200_crf_decode stays the default.  Same scores / outputs / workspace as b200_crf_decode
 * (the forward-backward pass runs first and provides the look-ahead scores and the qualities).
 */
int b200_crf_beam_search(const void* scores, int n, int t, int state_len, float blank_score, int beam_width, float beam_cut,
                         float qscale, float qbias, void* workspace, void* moves, void* sequence, void* qstring, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BONITO_B200_H */
