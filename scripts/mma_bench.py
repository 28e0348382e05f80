import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonito_b200 import native
iters = 960
for ts in (1, 0):
    for n, chain_list in ((32, (1, 2, 4, 6, 8, 12)), (64, (1, 2, 4, 6)), (128, (1, 2, 3))):
        for chains in chain_list:
            native.mma_bench(ts, n, iters, chains, 148)
            issue, total, ns = native.mma_bench(ts, n, iters, chains, 148)
            print(f"{'TS' if ts else 'SS'} N={n:3d} chains={chains:2d}: {total/iters:6.1f} cyc/MMA (issue {issue/iters:5.1f}), "
                  f"{128*n*16*2*iters/(ns*1e-9)/1e12*148:.0f} TFLOP/s chip")
