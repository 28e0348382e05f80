"""Hardware questions the recurrent-kernel design depends on (run on the B200 box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bonito_b200 import native
lib = native.require()
torch.zeros(1, device="cuda")
for cs in (1, 2, 4, 5, 6, 7, 8, 12, 16):
    for threads, smem in ((544, 120 * 1024), (800, 190 * 1024)):
        print(f"max active clusters: cluster_size={cs} threads={threads} smem={smem // 1024}K ->",
              lib.b200_debug_max_clusters(cs, threads, smem), flush=True)
print("tcgen05.mma M=128 K=16 fp16, A from TMEM: (issue cycles, issue-to-completion cycles, ns) per 960 MMAs")
for n in (16, 32, 48, 64, 96, 128):
    for chains in (1, 2, 3):
        if chains * n > 448:
            continue
        r = native.mma_bench(1, n, 960, chains, 1)
        print(f"  N={n:3d} chains={chains}: issue {r[0] / 960:.1f} cyc/MMA, done {r[1] / 960:.1f} cyc/MMA, {r[2]} ns", flush=True)
