"""One or two full-size steps (hac, batch 512, 9996 samples) for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda", 0)
model, spec, weights, chunksize = bench.build_model(dev, 0, 1)
from bonito_b200 import synth
from bonito_b200.decode import _decoder
x = synth.squiggle(64, chunksize, seed=100).repeat(batch // 64 + 1, 1, 1)[:batch].to(dev, torch.float16)
plan = model.native_plan(dev)
with torch.inference_mode():
    for _ in range(steps):
        s = plan.forward(x)
        _decoder(s, spec["state_len"], blank_score=2.0)
torch.cuda.synchronize()
print("done")
