"""Where the cycles of the weight-stationary GEMM go (B200_GEMM_DEBUG=1 counters, see gemm_tc.cu): per-CTA averages for
the hac input projection and CRF shapes at the headline batch."""
import os, sys
os.environ["B200_GEMM_DEBUG"] = str(1 | int(os.environ.get("GEMM_ABLATE", "0")))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import numpy as np
import torch
from bonito_b200 import native

T, TB, NT, H, CS, CW = 1666, 48, 11, 384, 6, 256
M = NT * T * TB


def profile(name, N, K, colblocks=False, act=native.ACT_NONE):
    a = (torch.randn(M, K, device="cuda") * 0.5).half()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    bias = torch.randn(N, device="cuda").half()
    c = torch.empty(M, N, dtype=torch.float16, device="cuda")
    kw = dict(rows_inner=TB, valid_inner=TB, stride_inner=1, stride_outer=CS * TB, cb_width=CW, cb_rows=TB) if colblocks else {}
    for _ in range(3):
        native.gemm(a, K, w, bias, c, CW if colblocks else N, M, N, K, act=act, lo=-5.0, hi=5.0, **kw)
    torch.cuda.synchronize()
    out = np.zeros(160 * 8, dtype=np.int64)
    rc = native.require().b200_debug_gemm_profile(ctypes.c_void_p(out.ctypes.data))
    assert rc == 0
    p = out.reshape(160, 8)
    p = p[p[:, 6] > 0]
    life, tiles = p[:, 0].mean(), p[:, 6].mean()
    print(f"[ablate {os.environ.get('GEMM_ABLATE', '0')}] {name}: {len(p)} CTAs, {tiles:.1f} tiles each, lifetime {life:.0f} cycles = {life / tiles:.0f} per tile")
    for i, label in ((1, "producer waits for a ring slot"), (2, "MMA thread waits for A"), (3, "MMA thread waits for an accumulator"),
                     (4, "epilogue warp waits for an accumulator"), (5, "epilogue warp inside epilogue_tile")):
        print(f"   {label:42s} {p[:, i].mean() / tiles:8.0f} cycles per tile  ({100 * p[:, i].mean() / life:5.1f} % of the lifetime)")


profile("lstm_in (N=1536, BN=192)", 4 * H, H, colblocks=True)
profile("crf (N=4096, BN=128)", 4096, H, act=native.ACT_CLAMP)
