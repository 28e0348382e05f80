"""Isolated correctness + timing of the hac GEMM shapes at the headline batch (512 chunks x 1666 steps) under the kernel
variant the environment selects (B200_GEMM_CLUSTER=<BN>x<CL> | 0, B200_GEMM_WS=0).  One process per variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bonito_b200 import native

T, TB, NT, H, CS, CW = 1666, 48, 11, 384, 6, 256
M = NT * T * TB


def bench(name, N, K, colblocks=False, act=native.ACT_NONE):
    g = torch.Generator(device="cuda").manual_seed(N + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g).half()
    c = torch.full((M, N), float("nan"), dtype=torch.float16, device="cuda")
    kw = dict(rows_inner=TB, valid_inner=TB, stride_inner=1, stride_outer=CS * TB, cb_width=CW, cb_rows=TB) if colblocks else {}
    ldc = CW if colblocks else N
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def run():
        native.gemm(a, K, w, bias, c, ldc, M, N, K, act=act, lo=-5.0, hi=5.0, **kw)

    run()
    torch.cuda.synchronize()
    # correctness on a sample of row blocks (first, last, a few in between) against fp32 matmul of the same fp16 operands
    got = c.view(NT * T, CS, TB, CW).permute(0, 2, 1, 3).reshape(M, N) if colblocks else c
    worst = 0.0
    for r0 in (0, 128 * 777, 128 * 3001, M - 300):
        ref = a[r0:r0 + 300].float() @ w.float().T + bias.float()
        if act == native.ACT_CLAMP:
            ref = ref.clamp(-5.0, 5.0)
        err = (got[r0:r0 + 300].float() - ref).abs().max().item()
        worst = max(worst, err)
    assert not torch.isnan(got).any().item(), "unwritten outputs"
    for _ in range(2):
        run()
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); t = ts[len(ts) // 2]
    print("%-8s M=%d N=%4d K=%3d  %.4f ms  %5.0f TFLOP/s   max|err| %.2e" % (name, M, N, K, t, 2.0 * M * N * K / t / 1e9, worst),
          flush=True)


print("env:", {k: v for k, v in os.environ.items() if k.startswith("B200_")}, flush=True)
bench("lstm_in", 4 * H, H, colblocks=True)
bench("crf", 4096, H, act=native.ACT_CLAMP)
bench("conv", H, 320)
