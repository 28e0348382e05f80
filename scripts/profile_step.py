"""One or two full-size steps for ncu captures: `profile_step.py [hac|sup] [steps] [batch]` (single stream, one batch in flight)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
which = sys.argv[1] if len(sys.argv) > 1 else "hac"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
from bonito_b200 import synth
from bonito_b200.decode import _decoder
if which == "hac":
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    model, spec, weights, chunksize = bench.build_hac(dev, 0, 1, batch=batch)
    x = synth.squiggle(64, chunksize, seed=100).repeat(batch // 64 + 1, 1, 1)[:batch].to(dev, torch.float16)
else:
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    model, spec = bench.build_sup(dev, 0, 1, batch=batch)
    x = synth.squiggle(32, 9996, seed=200).repeat(batch // 32 + 1, 1, 1)[:batch].to(dev, torch.float16)
plan = model.native_plan(dev)
with torch.inference_mode():
    for _ in range(steps):
        s = plan.forward(x)
        _decoder(s, spec["state_len"], blank_score=2.0)
torch.cuda.synchronize()
print("done")
