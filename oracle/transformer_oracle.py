"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the transformer (sup v5) forward path.

Follows `/root/reference/bonito/transformer/model.py:42-133` (MultiHeadAttention, TransformerEncoderLayer),
`bonito/nn.py:139-171` (LinearUpsample), `:268-298` (LinearCRFEncoder) and, for the arithmetic the reference delegates to
flash-attn (unpinned version, README.md:37-40), flash-attn's own torch reference functions:
`flash_attn/ops/triton/layer_norm.py:104-153` (rms_norm_ref: (x + residual) -> x * rsqrt(mean(x^2) + eps) * w, eps 1e-5),
`flash_attn/layers/rotary.py:14-36` (NeoX half rotation; cos/sin computed in fp32 then cast to the qkv dtype,
`:413-416`), `flash_attn/ops/activations.py:107-111` (swiglu: float(x)*float(y)/(1+exp(-x)), rounded once),
window semantics `flash_attn_interface.py:1016-1017` == `sliding_window_mask` (`transformer/model.py:33-39`).

Pinned (tests/golden/forward_sup.npz, written by oracle/make_golden.py) against the reference's own module classes
with the Triton / CUDA-only pieces swapped for those flash-attn reference functions and the SDPA branch of
`MultiHeadAttention.attn_func` (`transformer/model.py:61-65`); flash-attn's fused kernels themselves cannot run on
the CPU, so their rounding behaviour (documented in SURVEY.md Appendix A) is restated, not observed.
"""

import torch
import torch.nn.functional as F

from oracle.crf_oracle import _r16, convolution


def rotary_tables(seq_len, dim, fp16, base=10000.0):
    inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
    freqs = torch.outer(torch.arange(seq_len, dtype=torch.float32), inv_freq)
    return _r16(torch.cos(freqs), fp16), _r16(torch.sin(freqs), fp16)


def apply_rotary(x, cos, sin):
    """x [N, T, heads, dim]; rotate_half (non-interleaved)."""
    half = x.shape[-1] // 2
    x1, x2 = x[..., :half], x[..., half:]
    c, s = cos[None, :, None, :], sin[None, :, None, :]
    return torch.cat([x1 * c - x2 * s, x1 * s + x2 * c], dim=-1)


def window_mask(T, window):
    i = torch.arange(T)[:, None]
    j = torch.arange(T)[None, :]
    return (j >= i - window[0]) & (j <= i + window[1])


def rms_norm(x, residual, weight, eps=1e-5):
    s = x + residual
    return s * torch.rsqrt(s.square().mean(dim=-1, keepdim=True) + eps) * weight


def encoder_layer(x, w, prefix, nhead, alpha, window, fp16=False):
    """One TransformerEncoderLayer on x [N, T, d] (fp32 tensors holding fp16-representable values when fp16)."""
    N, T, d = x.shape
    hd = d // nhead
    qkv = _r16(F.linear(x, w[prefix + "self_attn.Wqkv.weight"]), fp16).view(N, T, 3, nhead, hd)
    cos, sin = rotary_tables(T, hd, fp16)
    q = _r16(apply_rotary(qkv[:, :, 0], cos, sin), fp16).permute(0, 2, 1, 3)
    k = _r16(apply_rotary(qkv[:, :, 1], cos, sin), fp16).permute(0, 2, 1, 3)
    v = qkv[:, :, 2].permute(0, 2, 1, 3)
    att = (q @ k.transpose(-1, -2)) / hd ** 0.5
    att = att.masked_fill(~window_mask(T, window), float("-inf"))
    o = _r16(torch.softmax(att, dim=-1) @ v, fp16).permute(0, 2, 1, 3).reshape(N, T, d)
    a = _r16(F.linear(o, w[prefix + "self_attn.out_proj.weight"], w[prefix + "self_attn.out_proj.bias"]), fp16)
    al = _r16(torch.tensor(alpha), fp16)
    x = _r16(rms_norm(a, _r16(al * x, fp16), w[prefix + "norm1.weight"]), fp16)
    h = _r16(F.linear(x, w[prefix + "ff.fc1.weight"]), fp16)
    y, gate = h.chunk(2, dim=-1)
    f = _r16(gate * y / (1.0 + torch.exp(-gate)), fp16)
    f = _r16(F.linear(f, w[prefix + "ff.fc2.weight"]), fp16)
    return _r16(rms_norm(f, _r16(al * x, fp16), w[prefix + "norm2.weight"]), fp16)


def transformer_forward(w, spec, x, fp16=False, return_features=False):
    """
    Whole sup encoder.  w: state-dict-named fp32 weights (BN already folded into conv weights);
    spec: dict(convs=[(cin,cout,k,stride,pad,act)...], depth, d_model, nhead, alpha, window, scale, state_len)
    x [N, 1, L] -> scores [N, 2T', C] (batch-first, no blank column).
    """
    feats = {}
    h = x
    for i, (_, _, _, stride, pad, act) in enumerate(spec["convs"]):
        h = convolution(h, w[f"conv.{i}.conv.weight"], w[f"conv.{i}.conv.bias"], stride, pad, act, fp16)
        feats[f"conv{i}"] = h
    h = h.permute(0, 2, 1)  # Permute([0,2,1]) -> [N, T', d]
    for l in range(spec["depth"]):
        h = encoder_layer(h, w, f"transformer_encoder.{l}.", spec["nhead"], spec["alpha"], spec["window"], fp16)
        feats[f"layer{l}"] = h
    N, T, d = h.shape
    up = _r16(F.linear(h, w["upsample.linear.weight"], w["upsample.linear.bias"]), fp16).reshape(N, 2 * T, d)
    feats["upsample"] = up
    s = _r16(F.linear(up, w["crf.linear.weight"]), fp16)
    s = _r16(s * spec["scale"], fp16)
    return (s, feats) if return_features else s
