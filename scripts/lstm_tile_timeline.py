"""Per-step timeline of the tile-layout tcgen05 LSTM kernel (CTA 0, sub-tile 0), hac batch 512 = 11 tiles of 48."""
import os, sys
os.environ["B200_LSTM_DEBUG"] = "3"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bonito_b200 import native
T, N, H = int(os.environ.get('TL_T', 400)), int(os.environ.get('TL_N', 512)), 384
tb, cs = native.lstm_tile_chunks(H), native.lstm_tile_cluster(H)
nt = -(-N // tb)
gx = (torch.randn(nt, T, cs, tb, 4 * H // cs, device="cuda") * 0.5).half()
whh = (torch.randn(4 * H, H, device="cuda") / H ** 0.5).half()
y = torch.empty(nt, T, tb, H, dtype=torch.float16, device="cuda")
for _ in range(int(os.environ.get('TL_REPS', 2))):
    native.lstm_rec_tile(gx, whh, y, T, N, H, False)
torch.cuda.synchronize()
for rev in (False, True):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    native.lstm_rec_tile(gx, whh, y, T, N, H, rev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("reverse=%s: %.3f ms for %d steps x %d tiles = %.2f us/step ; %.1f TFLOP/s on %d SMs" %
          (rev, ms, T, nt, ms * 1e3 / T, 2.0 * T * N * 4 * H * H / ms / 1e9, nt * cs))
native.lstm_rec_tile(gx, whh, y, T, N, H, False)
torch.cuda.synchronize()
tl = native.lstm_tile_timeline(256).astype(np.float64)
last = np.array([max(s for s in range(T) if s % 256 == i) for i in range(256)])
order = np.argsort(last)
tl = tl[order]
s = slice(20, 250)
names = ["h_full", "mma_issued", "d_full(w0)", "tmem_ld(w0)", "math(w0)", "sent(w0)", None, None]
print("cycles per step: %.0f" % np.diff(tl[s, 0]).mean())
base = tl[s, 0]
print("SM clock during the kernel: %.0f MHz" % ((tl[250, 0] - tl[20, 0]) / (tl[250, 6] - tl[20, 6]) * 1e3))
for i, n in enumerate(names):
    if n is None:
        continue
    print("%-12s +%7.0f cycles after h_full" % (n, (tl[s, i] - base).mean()))
print("next h_full after sent(w0): %.0f ; MMA warp waits for h_full(sub 0) for %.0f cycles" %
      ((tl[21:251, 0] - tl[20:250, 5]).mean(), (tl[s, 0] - tl[s, 7]).mean()))
