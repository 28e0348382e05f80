"""Per-tile timeline of the tcgen05 attention kernel (CTA 0), sup shape."""
import os, sys
os.environ["B200_ATTN_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import numpy as np, torch
from bonito_b200 import native
N, T, NH = int(os.environ.get("TL_N", 256)), 833, 8
qkv = (torch.randn(N, T, 3, NH, 64, device="cuda") * 1.5).half()
freqs = torch.outer(torch.arange(T, dtype=torch.float32), 1.0 / (10000.0 ** (torch.arange(0, 64, 2, dtype=torch.float32) / 64)))
cos_sin = torch.cat([torch.cos(freqs), torch.sin(freqs)], dim=1).half().cuda()
out = torch.empty(N, T, NH * 64, dtype=torch.float16, device="cuda")
for _ in range(2):
    native.attention(qkv, cos_sin, out, N, T, NH, 64, 127, 128)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); native.attention(qkv, cos_sin, out, N, T, NH, 64, 127, 128); e1.record(); torch.cuda.synchronize()
print("rotary + attention: %.3f ms" % e0.elapsed_time(e1))
buf = np.zeros((64, 16), dtype=np.int64)
n = native.load().b200_debug_attention_timeline(buf.ctypes.data_as(ctypes.c_void_p), 64)
tl = buf[:n].astype(np.float64)
if os.environ.get("B200_ATTN_TC") == "1":
    names = ["mma: Q+tmem ready", "mma: QK issued", "mma: PV issued", "mma: O complete", "mma: loads issued", "sm: s(0) seen", "sm: S0 loaded",
             "sm: max exchanged", "sm: P0 arrived", "sm: s(1) seen", "sm: P1 arrived", "sm: P2 arrived", "sm: o seen", "sm: O loaded", "sm: out stored"]
else:    # second version: one pass for the maxima, one for the probabilities
    names = ["mma: Q ready", "mma: QK issued", "mma: PV issued", "mma: O complete", "mma: loads issued", "sm: s(0) seen", "sm: pass 1 done",
             "sm: max exchanged", "sm: P arrived", None, None, None, "sm: o seen", "sm: O loaded", "sm: out stored"]
sel = slice(10, 50)
base = tl[sel, 0]
print("cycles per tile: %.0f" % np.diff(tl[sel, 0]).mean())
for i, nm in enumerate(names):
    if nm is None:
        continue
    print("%-22s +%7.0f" % (nm, (tl[sel, i] - base).mean()))
