"""Isolated attention comparison on the same box: this repo's rotary + windowed attention (b200_attention_fwd, tcgen05 and
mma.sync kernels) against flash-attn's `flash_attn_qkvpacked_func(window_size=(127, 128))` (what the reference's
MultiHeadAttention.attn_func calls, bonito/transformer/model.py:55-60; rotary applied by flash-attn's own kernel before).
sup shape: batch 256, 833 tokens, 8 heads of 64.  CUDA events, 5 warm-up + 20 timed launches each."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bonito_b200 import native
N, T, NH = 256, 833, 8
g = torch.Generator(device="cuda").manual_seed(1)
qkv0 = (torch.randn(N, T, 3, NH, 64, device="cuda", generator=g) * 1.5).half()
inv_freq = 1.0 / (10000.0 ** (torch.arange(0, 64, 2, dtype=torch.float32, device="cuda") / 64))
freqs = torch.outer(torch.arange(T, dtype=torch.float32, device="cuda"), inv_freq)
cos_sin = torch.cat([torch.cos(freqs), torch.sin(freqs)], dim=1).half()
out = torch.empty(N, T, NH * 64, dtype=torch.float16, device="cuda")


def timed(fn, reps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {}
qkv = qkv0.clone()
os.environ.pop("B200_ATTN_IMPL", None)
res["b200 tcgen05 (rotary + attention) ms"] = timed(lambda: native.attention(qkv, cos_sin, out, N, T, NH, 64, 127, 128))
ours = out.clone()
os.environ["B200_ATTN_IMPL"] = "mma"
res["b200 mma.sync (rotary + attention) ms"] = timed(lambda: native.attention(qkv, cos_sin, out, N, T, NH, 64, 127, 128))
os.environ.pop("B200_ATTN_IMPL", None)
try:
    from flash_attn import flash_attn_qkvpacked_func
    from flash_attn.layers.rotary import RotaryEmbedding
    rot = RotaryEmbedding(64, interleaved=False).cuda()
    q2 = qkv0.clone()
    res["flash-attn 2 rotary + qkvpacked(window 127/128) ms"] = timed(
        lambda: flash_attn_qkvpacked_func(rot(q2.clone()), window_size=(127, 128)))
    res["flash-attn 2 qkvpacked(window 127/128) only ms"] = timed(lambda: flash_attn_qkvpacked_func(q2, window_size=(127, 128)))
    import flash_attn
    res["flash_attn version"] = flash_attn.__version__
    # numerics: same input, rotary by flash-attn, attention by flash-attn vs ours
    qq = qkv0.clone()
    native.attention(qq, cos_sin, out, N, T, NH, 64, 127, 128)
    ref = flash_attn_qkvpacked_func(rot(qkv0.clone()), window_size=(127, 128)).reshape(N, T, NH * 64)
    res["max |b200 - flash-attn|"] = (out.float() - ref.float()).abs().max().item()
except Exception as err:  # flash-attn not importable on this box
    res["flash-attn"] = f"unavailable: {err!r}"
res["speed-up tcgen05 vs flash-attn (rotary included)"] = (
    res.get("flash-attn 2 rotary + qkvpacked(window 127/128) ms", float("nan")) / res["b200 tcgen05 (rotary + attention) ms"])
print(json.dumps(res, indent=1))
