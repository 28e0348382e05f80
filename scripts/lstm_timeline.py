"""Per-step timeline of the tcgen05 LSTM kernel (CTA 0), hac batch 512."""
import os, sys
os.environ["B200_LSTM_DEBUG"] = "3"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bonito_b200 import native
T, N, H = int(os.environ.get('TL_T', 400)), 512, 384
gx = (torch.randn(T, N, 4 * H, device="cuda") * 0.5).half()
whh = (torch.randn(4 * H, H, device="cuda") / H ** 0.5).half()
y = torch.empty(T, N, H, dtype=torch.float16, device="cuda")
for _ in range(int(os.environ.get('TL_REPS', 2))):
    native.lstm_rec(gx, whh, y, T, N, H, False)
torch.cuda.synchronize()
for rev in (False, True):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    native.lstm_rec(gx, whh, y, T, N, H, rev)
    e1.record()
    torch.cuda.synchronize()
    print("reverse=%s: %.3f ms for %d steps = %.2f us/step" % (rev, e0.elapsed_time(e1), T, e0.elapsed_time(e1) * 1e3 / T))
native.lstm_rec(gx, whh, y, T, N, H, False)
torch.cuda.synchronize()
tl = native.lstm_timeline(256).astype(np.float64)
# ring buffer: slot i holds the last step s < T with s % 256 == i; put back in step order
last = np.array([max(s for s in range(T) if s % 256 == i) for i in range(256)])
order = np.argsort(last)
tl = tl[order]
print("steps recorded: %d..%d" % (last.min(), last.max()))
s = slice(20, 250)
names = ["h_full", "mma_issued", "d_full(w0)", "tmem_ld(w0)", "math(w0)", "sent(w0)", None, "sent(w7)"]
step = np.diff(tl[s, 0]).mean()
print("cycles per step: %.0f" % step)
base = tl[s, 0]
print("SM clock during the kernel: %.0f MHz" % ((tl[250, 0] - tl[20, 0]) / (tl[250, 6] - tl[20, 6]) * 1e3))
for i, n in enumerate(names):
    if n is None:
        continue
    print("%-12s +%7.0f cycles after h_full" % (n, (tl[s, i] - base).mean()))
nxt = tl[21:251, 0] - tl[20:250, 5]
print("next h_full after sent(w0): %.0f ; after sent(w7): %.0f" % (nxt.mean(), (tl[21:251, 0] - tl[20:250, 7]).mean()))
