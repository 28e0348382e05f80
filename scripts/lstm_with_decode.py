"""Does a co-resident CRF decode slow the recurrent kernel down?  15 clusters (480 chunks) + decode of 512 chunks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bonito_b200 import native
T, N, H = 1666, 480, 384
gx = (torch.randn(T, N, 4 * H, device="cuda") * 0.5).half()
whh = (torch.randn(4 * H, H, device="cuda") / H ** 0.5).half()
y = torch.empty(T, N, H, dtype=torch.float16, device="cuda")
sc = (torch.randn(512, T, 1024, device="cuda") * 1.5).clamp(-5, 5).half()
ws = torch.empty(native.crf_decode_workspace_bytes(512, T, 4), dtype=torch.uint8, device="cuda")
outs = [torch.empty(512, T, dtype=torch.uint8, device="cuda") for _ in range(3)]
hi, lo = torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=0)
def ev(): return torch.cuda.Event(enable_timing=True)
def lstm(st): native.lstm_rec(gx, whh, y, T, N, H, False, stream=st)
def dec(st): native.crf_decode(sc, 4, 2.0, 1.0, 0.0, ws, *outs, stream=st)
for _ in range(2): lstm(hi); dec(lo)
torch.cuda.synchronize()
def run(order):
    a0, a1, b0, b1, t0, t1 = (ev() for _ in range(6))
    t0.record()
    hi.wait_event(t0); lo.wait_event(t0)
    for what in order:
        if what == "L": a0.record(hi); lstm(hi); a1.record(hi)
        else: b0.record(lo); dec(lo); b1.record(lo)
    torch.cuda.current_stream().wait_stream(hi); torch.cuda.current_stream().wait_stream(lo)
    t1.record(); torch.cuda.synchronize()
    return (a0.elapsed_time(a1) if "L" in order else 0, b0.elapsed_time(b1) if "D" in order else 0, t0.elapsed_time(t1))
for order in ("L", "D", "LD", "DL", "LD", "DL"):
    print(order, "lstm %.2f ms  decode %.2f ms  total %.2f ms" % run(order))
