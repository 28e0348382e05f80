"""Isolated timing of the three hac GEMM shapes (one 32-chunk tile) under the kernel variants selected by the environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bonito_b200 import native
M = 1666 * 32
def bench(name, N, K, remap=False, act=native.ACT_NONE):
    a = (torch.randn(M, K, device="cuda") * 0.5).half()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    bias = torch.randn(N, device="cuda").half()
    c = torch.empty(M, N, dtype=torch.float16, device="cuda")
    kw = dict(rows_inner=32, valid_inner=32, stride_inner=1666, stride_outer=1) if remap else {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    def run(): native.gemm(a, K, w, bias, c, N, M, N, K, act=act, lo=-5.0, hi=5.0, **kw)
    for _ in range(3): run()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); t = ts[len(ts) // 2]
    print("%-10s M=%d N=%4d K=%3d  %.4f ms  %.0f TFLOP/s" % (name, M, N, K, t, 2.0 * M * N * K / t / 1e9))
print("env:", {k: v for k, v in os.environ.items() if k.startswith("B200_")})
bench("conv", 384, 320)
bench("lstm_in", 1536, 384)
bench("crf", 1024, 384, remap=True, act=native.ACT_CLAMP)
