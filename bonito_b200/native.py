"""
ctypes binding of `libbonito_b200.so` (C ABI: include/bonito_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C bonito_b200/csrc`.
There is no CPU fallback: if the library (or a CUDA device) is missing, the native
path raises -- see `require()`.
"""

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p

import torch

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libbonito_b200.so")
_lib = None

ACT_NONE, ACT_SWISH, ACT_TANH, ACT_CLAMP, ACT_SCALE, ACT_SWIGLU, ACT_TANH_SCALE = 0, 1, 2, 3, 4, 5, 6
GEMM_AUTO, GEMM_TCGEN05, GEMM_MMA_SYNC, GEMM_TCGEN05_PAIR = 0, 1, 2, 3

MAX_LSTM_LAYERS = 8


class LstmCrfPlanStruct(ctypes.Structure):
    """`b200_lstm_crf_plan` of include/bonito_b200.h."""
    _fields_ = [("n", c_int), ("l", c_int), ("t", c_int), ("tp", c_int),
                ("c1", c_int), ("k1", c_int), ("act1", c_int), ("c2", c_int), ("k2", c_int), ("act2", c_int),
                ("hidden", c_int), ("k3", c_int), ("s3", c_int), ("pad3", c_int), ("act3", c_int),
                ("n_lstm", c_int), ("n_scores", c_int), ("act_l", c_int),
                ("lo", c_float), ("hi", c_float),
                ("reverse", c_int * MAX_LSTM_LAYERS),
                ("w1", c_void_p), ("b1", c_void_p), ("w2", c_void_p), ("b2", c_void_p), ("w3", c_void_p), ("b3", c_void_p),
                ("wl", c_void_p), ("bl", c_void_p),
                ("wih", c_void_p * MAX_LSTM_LAYERS), ("bias", c_void_p * MAX_LSTM_LAYERS), ("whh", c_void_p * MAX_LSTM_LAYERS),
                ("stem", c_void_p), ("ya", c_void_p), ("yb", c_void_p), ("gx", c_void_p), ("hx", c_void_p)]


# name -> (restype, argtypes); must list every symbol declared in include/bonito_b200.h
SIGNATURES = {
    "b200_version": (c_int, []),
    "b200_last_error": (c_char_p, []),
    "b200_conv_stem_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                   c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "b200_gemm_fwd": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_int,
                              c_int, c_float, c_float, c_int, c_int, c_longlong, c_longlong, c_int, c_void_p]),
    "b200_gemm_fwd_ex": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_int,
                                 c_int, c_float, c_float, c_int, c_int, c_longlong, c_longlong, c_int, c_longlong, c_int, c_int,
                                 c_int, c_int, c_void_p]),
    "b200_conv_first_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                    c_int, c_void_p]),
    "b200_attention_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200_rmsnorm_residual_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p, c_longlong, c_int,
                                          c_void_p]),
    "b200_swiglu_fwd": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_void_p]),
    "b200_lstm_cluster_size": (c_int, [c_int]),
    "b200_lstm_rec_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200_lstm_tile_chunks": (c_int, [c_int]),
    "b200_lstm_tile_cluster": (c_int, [c_int]),
    "b200_lstm_rec_tile_workspace_bytes": (c_size_t, [c_int]),
    "b200_lstm_rec_tile_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200_debug_lstm_tile_timeline": (c_int, [c_void_p, c_int]),
    "b200_debug_attention_timeline": (c_int, [c_void_p, c_int]),
    "b200_debug_gemm_profile": (c_int, [c_void_p]),
    "b200_debug_tmem_probe": (c_int, [c_void_p, c_void_p]),
    "b200_debug_lstm_timeline": (c_int, [c_void_p, c_int]),
    "b200_debug_lstm_max_clusters": (c_int, []),
    "b200_debug_max_clusters": (c_int, [c_int, c_int, c_int]),
    "b200_debug_exchange_bench": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "b200_debug_mma_bench": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "b200_crf_decode_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "b200_quantize_i8": (c_int, [c_void_p, c_void_p, c_longlong, c_float, c_void_p]),
    "b200_stream_create": (c_int, [c_void_p]),
    "b200_chunk_count": (c_int, [c_longlong, c_int, c_int]),
    "b200_chunk_signal": (c_int, [c_void_p, c_int, c_longlong, c_int, c_int, c_void_p, c_longlong, c_void_p]),
    "b200_gemm_i8_fwd": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_int,
                                 c_int, c_float, c_float, c_int, c_int, c_longlong, c_longlong, c_int, c_longlong, c_int, c_int,
                                 c_int, c_void_p]),
    "b200_lstm_crf_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200_crf_beam_search": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_int, c_float, c_float, c_float,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200_crf_decode": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_float,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}


class NativeError(RuntimeError):
    pass


def lib_path():
    return _LIB_PATH


def load():
    """Load the shared library (no CUDA calls are made by loading it)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise NativeError(
                f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C bonito_b200/csrc` (there is no CPU fallback for the native path)")
        lib = ctypes.CDLL(_LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = restype, argtypes
        _lib = lib
    return _lib


def require(device=None):
    """Library + CUDA device, or a loud failure."""
    lib = load()
    if not torch.cuda.is_available():
        raise NativeError("bonito_b200 native path needs a CUDA device (sm_100a); none is visible")
    return lib


def _check(rc, what):
    if rc != 0:
        msg = load().b200_last_error().decode(errors="replace")
        raise NativeError(f"{what} failed ({rc}): {msg}")


def _ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def _stream(stream=None):
    """cudaStream_t for the C ABI: an explicit torch stream, or the current one."""
    return c_void_p((stream or torch.cuda.current_stream()).cuda_stream)


def _f16(t, name):
    if t.dtype != torch.float16 or not t.is_cuda or not t.is_contiguous():
        raise NativeError(f"{name}: expected a contiguous CUDA fp16 tensor, got {t.dtype} {t.device}")
    return t


def version():
    return load().b200_version()


def conv_stem(x, w1, b1, act1, w2, b2, act2, out, lp, padl):
    """x [N,L] -> out [N,lp,C2] channels-last, zero padded (see b200_conv_stem_fwd)."""
    lib = require()
    n, l = x.shape
    c1, _, k1 = w1.shape
    c2, _, k2 = w2.shape
    with torch.cuda.device(x.device):
        rc = lib.b200_conv_stem_fwd(_ptr(_f16(x, "x")), n, l, c1, k1, _ptr(_f16(w1, "w1")), _ptr(b1), act1,
                                    c2, k2, _ptr(_f16(w2, "w2")), _ptr(b2), act2, _ptr(out), lp, padl, _stream())
    _check(rc, "b200_conv_stem_fwd")
    return out


def gemm(a_ptr_tensor, lda, b, bias, c, ldc, m, n, k, act=ACT_NONE, lo=0.0, hi=0.0,
         rows_inner=None, valid_inner=None, stride_inner=1, stride_outer=0, impl=GEMM_AUTO, stream=None, max_ctas=0,
         cb_width=0, cb_rows=0, group=0, stride_group=0):
    """C = act(A B^T + bias); `a_ptr_tensor` / `c` only supply base pointers (rows may overlap / be remapped; `cb_*`:
    column blocks, see b200_gemm_fwd_ex)."""
    lib = require()
    if rows_inner is None:
        rows_inner, valid_inner = m, m
    with torch.cuda.device(c.device):
        rc = lib.b200_gemm_fwd_ex(_ptr(a_ptr_tensor), lda, _ptr(_f16(b, "b")), _ptr(bias), _ptr(c), ldc, m, n, k,
                                  act, float(lo), float(hi), rows_inner, valid_inner, stride_inner, stride_outer,
                                  int(group), int(stride_group), int(cb_width), int(cb_rows), impl, int(max_ctas),
                                  _stream(stream))
    _check(rc, "b200_gemm_fwd")
    return c


def conv_first(x, w, bias, act, out, lp, padl, stream=None):
    """x [N,L] -> out [N,lp,C] channels-last with zero halo (see b200_conv_first_fwd)."""
    lib = require()
    n, l = x.shape
    c, _, k = w.shape
    with torch.cuda.device(out.device):
        rc = lib.b200_conv_first_fwd(_ptr(_f16(x, "x")), n, l, c, k, _ptr(_f16(w, "w")), _ptr(bias), act, _ptr(out), lp,
                                     padl, _stream(stream))
    _check(rc, "b200_conv_first_fwd")
    return out


def attention(qkv, cos_sin, out, n, t, heads, head_dim, wl, wr, stream=None):
    lib = require()
    with torch.cuda.device(out.device):
        rc = lib.b200_attention_fwd(_ptr(_f16(qkv, "qkv")), _ptr(_f16(cos_sin, "cos_sin")), _ptr(out), n, t, heads,
                                    head_dim, wl, wr, _stream(stream))
    _check(rc, "b200_attention_fwd")
    return out


def rmsnorm_residual(a, x, w, alpha, eps, out, m, d, stream=None):
    lib = require()
    with torch.cuda.device(out.device):
        rc = lib.b200_rmsnorm_residual_fwd(_ptr(_f16(a, "a")), _ptr(_f16(x, "x")), _ptr(_f16(w, "w")), float(alpha),
                                           float(eps), _ptr(out), m, d, _stream(stream))
    _check(rc, "b200_rmsnorm_residual_fwd")
    return out


def swiglu(h, out, m, f, stream=None):
    lib = require()
    with torch.cuda.device(out.device):
        rc = lib.b200_swiglu_fwd(_ptr(_f16(h, "h")), _ptr(out), m, f, _stream(stream))
    _check(rc, "b200_swiglu_fwd")
    return out


def lstm_cluster_size(hidden):
    return load().b200_lstm_cluster_size(hidden)


def lstm_rec(gx, whh, y, t, n, hidden, reverse, stream=None):
    lib = require()
    with torch.cuda.device(y.device):
        rc = lib.b200_lstm_rec_fwd(_ptr(gx), _ptr(_f16(whh, "whh")), _ptr(y), t, n, hidden,
                                   int(bool(reverse)), _stream(stream))
    _check(rc, "b200_lstm_rec_fwd")
    return y


def lstm_tile_chunks(hidden):
    """Chunks per tile of the tile-layout recurrent kernel (0: this hidden size only has the generic-layout kernel)."""
    return load().b200_lstm_tile_chunks(hidden)


def lstm_tile_cluster(hidden):
    return load().b200_lstm_tile_cluster(hidden)


def lstm_rec_tile_workspace_bytes(n):
    return load().b200_lstm_rec_tile_workspace_bytes(n)


def lstm_rec_tile(gx, whh, y, t, n, hidden, reverse, stream=None, workspace=None):
    """gx [tiles][T][6][48][256], y [tiles][T][48][H] (see b200_lstm_rec_tile_fwd); n chunks = ceil(n/48) tiles.
    `workspace`: uint8 tensor of lstm_rec_tile_workspace_bytes(n) bytes (allocated here when omitted)."""
    lib = require()
    if workspace is None:
        workspace = torch.empty(lstm_rec_tile_workspace_bytes(n), dtype=torch.uint8, device=y.device)
    with torch.cuda.device(y.device):
        rc = lib.b200_lstm_rec_tile_fwd(_ptr(gx), _ptr(_f16(whh, "whh")), _ptr(y), _ptr(workspace), t, n, hidden,
                                        int(bool(reverse)), _stream(stream))
    _check(rc, "b200_lstm_rec_tile_fwd")
    return y


def lstm_tile_timeline(steps=256):
    """[steps, 8] int64 SM-clock stamps recorded by CTA 0 of the last lstm_rec_tile launch under B200_LSTM_DEBUG=3."""
    import numpy as np
    buf = np.zeros((steps, 8), dtype=np.int64)
    n = load().b200_debug_lstm_tile_timeline(buf.ctypes.data_as(c_void_p), steps)
    if n < 0:
        _check(n, "b200_debug_lstm_tile_timeline")
    return buf[:n]


def tmem_probe():
    """Run the TMEM convention probe; returns a float32 CPU tensor of 16384 values."""
    lib = require()
    out = torch.zeros(16384, dtype=torch.float32, device="cuda")
    rc = lib.b200_debug_tmem_probe(_ptr(out), _stream())
    _check(rc, "b200_debug_tmem_probe")
    torch.cuda.synchronize()
    return out.cpu()


def mma_bench(ts_mode, n, iters=960, chains=1, blocks=1):
    """(issue cycles, issue-to-completion cycles, ns) of tcgen05.mma M=128 x N=n x K=16 over `chains` accumulators."""
    lib = require()
    out = torch.zeros(3, dtype=torch.int64, device="cuda")
    _check(lib.b200_debug_mma_bench(int(ts_mode), n, iters, chains, blocks, _ptr(out), _stream()), "b200_debug_mma_bench")
    torch.cuda.synchronize()
    return out.cpu().tolist()


def lstm_timeline(steps=256):
    """[steps, 8] int64 SM-clock stamps recorded by CTA 0 of the last lstm_rec launch under B200_LSTM_DEBUG=3."""
    import numpy as np
    buf = np.zeros((steps, 8), dtype=np.int64)
    n = load().b200_debug_lstm_timeline(buf.ctypes.data_as(c_void_p), steps)
    if n < 0:
        _check(n, "b200_debug_lstm_timeline")
    return buf[:n]


def crf_decode_workspace_bytes(n, t, state_len):
    return load().b200_crf_decode_workspace_bytes(n, t, state_len)


def crf_decode(scores, state_len, blank_score, qscale, qbias, workspace, moves, sequence, qstring, stream=None):
    lib = require()
    n, t, _ = scores.shape
    with torch.cuda.device(scores.device):
        rc = lib.b200_crf_decode(_ptr(scores), n, t, state_len, float(blank_score), float(qscale),
                                 float(qbias), _ptr(workspace), _ptr(moves), _ptr(sequence), _ptr(qstring),
                                 _stream(stream))
    _check(rc, "b200_crf_decode")
    return moves, sequence, qstring


def crf_beam_search(scores, state_len, blank_score, beam_width, beam_cut, qscale, qbias, workspace, moves, sequence, qstring,
                    stream=None):
    lib = require()
    n, t, _ = scores.shape
    with torch.cuda.device(scores.device):
        rc = lib.b200_crf_beam_search(_ptr(scores), n, t, state_len, float(blank_score), int(beam_width), float(beam_cut),
                                      float(qscale), float(qbias), _ptr(workspace), _ptr(moves), _ptr(sequence),
                                      _ptr(qstring), _stream(stream))
    _check(rc, "b200_crf_beam_search")
    return moves, sequence, qstring


def lstm_crf_fwd(plan_struct, x, scores, stream=None):
    """Whole encoder forward from one C call (see b200_lstm_crf_fwd); `plan_struct`: a filled LstmCrfPlanStruct."""
    lib = require()
    with torch.cuda.device(scores.device):
        rc = lib.b200_lstm_crf_fwd(ctypes.byref(plan_struct), _ptr(_f16(x, "x")), _ptr(scores), _stream(stream))
    _check(rc, "b200_lstm_crf_fwd")
    return scores


def chunk_signal(signal, chunksize, overlap, out=None, stream=None):
    """bonito.util.chunk for ONE read already on the device: signal [length] (or [1, length]) fp16 / fp32 ->
    [n_chunks, 1, chunksize] fp16 from one gather kernel (see b200_chunk_signal)."""
    lib = require()
    sig = signal.reshape(-1)
    if not sig.is_cuda or sig.dtype not in (torch.float16, torch.float32) or not sig.is_contiguous():
        raise ValueError("chunk_signal: a contiguous float16 / float32 CUDA tensor expected")
    n = lib.b200_chunk_count(sig.numel(), int(chunksize), int(overlap))
    if n <= 0:
        raise ValueError(f"chunk_signal: bad geometry (length {sig.numel()}, chunksize {chunksize}, overlap {overlap})")
    if out is None:
        out = torch.empty(n, 1, chunksize, dtype=torch.float16, device=sig.device)
    with torch.cuda.device(sig.device):
        rc = lib.b200_chunk_signal(_ptr(sig), int(sig.dtype == torch.float32), sig.numel(), int(chunksize), int(overlap),
                                   _ptr(out), int(chunksize), _stream(stream))
    _check(rc, "b200_chunk_signal")
    return out


def new_stream(device):
    """A CUDA stream of its own (cudaStreamCreateWithFlags, non-blocking) wrapped for torch.  torch.cuda.Stream() returns
    one of 32 pooled streams per device round-robin, so streams that must run concurrently can silently be the same stream."""
    lib = require()
    handle = c_void_p()
    with torch.cuda.device(device):
        rc = lib.b200_stream_create(ctypes.byref(handle))
    _check(rc, "b200_stream_create")
    return torch.cuda.ExternalStream(handle.value, device=device)


def quantize_i8(x, out, scale=127.0, stream=None):
    """fp16 -> int8 (see b200_quantize_i8); `out`: int8 tensor with x.numel() elements."""
    lib = require()
    with torch.cuda.device(out.device):
        rc = lib.b200_quantize_i8(_ptr(_f16(x, "x")), _ptr(out), x.numel(), float(scale), _stream(stream))
    _check(rc, "b200_quantize_i8")
    return out


def gemm_i8(a, lda, b, col_scale, bias, c, ldc, m, n, k, act=ACT_NONE, lo=0.0, hi=0.0, rows_inner=None, valid_inner=None,
            stride_inner=1, stride_outer=0, group=0, stride_group=0, cb_width=0, cb_rows=0, stream=None, max_ctas=0):
    """C = act(col_scale * (A_i8 B_i8^T) + bias) (see b200_gemm_i8_fwd)."""
    lib = require()
    if rows_inner is None:
        rows_inner, valid_inner = m, m
    with torch.cuda.device(c.device):
        rc = lib.b200_gemm_i8_fwd(_ptr(a), lda, _ptr(b), _ptr(col_scale), _ptr(bias), _ptr(c), ldc, m, n, k, act, float(lo),
                                  float(hi), rows_inner, valid_inner, stride_inner, stride_outer, int(group), int(stride_group),
                                  int(cb_width), int(cb_rows), int(max_ctas), _stream(stream))
    _check(rc, "b200_gemm_i8_fwd")
    return c
