"""h all-gather micro-benchmark (see b200_debug_exchange_bench): cycles per step by mechanism and fake-compute delay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bonito_b200 import native
lib = native.require()
steps = 400
out = torch.zeros(2, dtype=torch.int64, device="cuda")
names = {0: "DSMEM bulk 256 B x 6 peers / warp", 2: "L2 staging + multicast bulk 256 B / warp",
         3: "L2 staging + multicast bulk 2 KB / sub-tile", 4: "DSMEM bulk 2 KB x 6 peers / sub-tile"}
for clusters in (1, 11):
    staging = torch.zeros(clusters * 73728, dtype=torch.uint8, device="cuda")
    for mode in (0, 4, 2, 3):
        for delay in (0, 800, 1600):
            rc = lib.b200_debug_exchange_bench(mode, steps, delay, clusters, staging.data_ptr(), out.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream)
            assert rc == 0, lib.b200_last_error()
            torch.cuda.synchronize()
            c, n = out.tolist()
            print(f"clusters={clusters:2d} mode {mode} ({names[mode]:45s}) delay {delay:5d}: {c / n:8.0f} cycles/step "
                  f"(exchange share {c / n - delay:6.0f})", flush=True)
