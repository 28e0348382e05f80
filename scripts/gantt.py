"""Launch-level timeline of one tile-pipelined step (start/end of every kernel relative to the step start)."""
import sys, os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", os.environ.get("CONN", "32"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from bonito_b200 import synth
from bonito_b200.decode import _decoder
dev = torch.device("cuda", 0)
model, spec, weights, chunksize = bench.build_model(dev, 0, 1)
x = synth.squiggle(64, chunksize, seed=100).repeat(8, 1, 1).to(dev, torch.float16)
plan = model.native_plan(dev)
with torch.inference_mode():
    for _ in range(3):
        _decoder(plan.forward(x), 4)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t0.record()
    ev = []
    s = plan.forward(x, events=ev)
    _decoder(s, 4, events=ev)
    t1 = torch.cuda.Event(enable_timing=True); t1.record()
torch.cuda.synchronize()
print("step %.2f ms" % t0.elapsed_time(t1))
rows = [(n, t0.elapsed_time(a), t0.elapsed_time(b)) for n, a, b in ev]
for name in ("conv_stem", "conv_gemm", "lstm_in_gemm", "lstm_rec", "crf_gemm", "crf_decode"):
    r = [(a, b) for n, a, b in rows if n == name]
    if not r: continue
    d = [b - a for a, b in r]
    print("%-13s n=%3d first start %6.2f last end %6.2f  dur min/avg/max %.2f/%.2f/%.2f" % (
        name, len(r), min(a for a, _ in r), max(b for _, b in r), min(d), sum(d) / len(d), max(d)))
# concurrency of lstm_rec over time
import numpy as np
r = [(a, b) for n, a, b in rows if n == "lstm_rec"]
grid = np.arange(0, max(b for _, b in r), 0.25)
conc = [sum(1 for a, b in r if a <= t < b) for t in grid]
print("lstm_rec launches in flight (every 0.25 ms):", "".join("%x" % min(c, 15) for c in conc))
# per-layer: lstm_rec launches are appended layer-major (16 tiles per layer)
for layer in range(5):
    rl = r[layer * 16:(layer + 1) * 16]
    print("layer %d: starts %s" % (layer, " ".join("%.1f" % a for a, _ in rl)))
    print("         ends   %s" % (" ".join("%.1f" % b for _, b in rl)))
