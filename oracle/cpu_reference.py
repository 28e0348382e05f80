"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  The reference's PyTorch-CPU execution of the path, rebuilt from torch
modules exactly as `bonito.nn` builds them (torch.nn.Conv1d / torch.nn.LSTM with flip / torch.nn.Linear / clamp:
`/root/reference/bonito/nn.py:226,361-370,279-298,66-67`), followed by the oracle's restatement of
`SeqdistModel.decode_batch` (the reference's basecaller decode is CUDA-only koi, `bonito/crf/basecall.py:36-40`): the
OpenMP C version (`oracle/csrc/crf_decode_ref.c`, bit-identical to the numpy one on the fixtures) so that the CPU arm is not
handicapped by numpy overheads.

Used by `bench.py` for the `cpu_baseline` leg and the `--impl reference` arm (the reference itself is Python and
cannot travel to the GPU box; its sources are never copied).
"""

import time

import numpy as np
import torch

from oracle import build_ref, crf_oracle


class CpuReferenceModel(torch.nn.Module):
    def __init__(self, spec, weights):
        super().__init__()
        self.spec = spec
        self.convs = torch.nn.ModuleList()
        self.acts = []
        for i, (cin, cout, k, stride, pad, act) in enumerate(spec["convs"]):
            conv = torch.nn.Conv1d(cin, cout, k, stride=stride, padding=pad, bias=True)
            conv.weight.data.copy_(weights[f"conv{i}.weight"])
            conv.bias.data.copy_(weights[f"conv{i}.bias"])
            self.convs.append(conv)
            self.acts.append({"swish": torch.nn.functional.silu, "tanh": torch.tanh, None: lambda v: v}[act])
        H = spec["hidden"]
        self.lstms = torch.nn.ModuleList()
        for i in range(spec["n_lstm"]):
            rnn = torch.nn.LSTM(H, H)
            rnn.weight_ih_l0.data.copy_(weights[f"lstm{i}.w_ih"])
            rnn.weight_hh_l0.data.copy_(weights[f"lstm{i}.w_hh"])
            rnn.bias_ih_l0.data.copy_(weights[f"lstm{i}.b_ih"])
            rnn.bias_hh_l0.data.copy_(weights[f"lstm{i}.b_hh"])
            self.lstms.append(rnn)
        self.linear = torch.nn.Linear(H, weights["crf.weight"].shape[0], bias=False)
        self.linear.weight.data.copy_(weights["crf.weight"])
        self.eval()

    @torch.inference_mode()
    def forward(self, x):
        """x [N,1,L] fp32 -> scores [T,N,C] (no blank column)."""
        h = x
        for conv, act in zip(self.convs, self.acts):
            h = act(conv(h))
        h = h.permute(2, 0, 1)
        for rnn, rev in zip(self.lstms, self.spec["reverse"]):
            if rev:
                h = h.flip(0)
            h, _ = rnn(h)
            if rev:
                h = h.flip(0)
        s = self.linear(h)
        if self.spec.get("clamp") is not None:
            s = s.clamp(*self.spec["clamp"])
        return s

    def basecall_batch(self, x, decode="c", decode_slice=4):
        """forward + decode; returns (moves, seq, qual, t_forward, t_decode).  decode: "c" (OpenMP) or "numpy"."""
        t0 = time.perf_counter()
        s = self.forward(x)
        t1 = time.perf_counter()
        ntc = s.permute(1, 0, 2).contiguous().numpy()
        if decode == "c":
            moves, seq, qual = build_ref.decode(ntc, self.spec["state_len"], self.spec["blank_score"])
        else:
            outs = [crf_oracle.decode_native(ntc[i:i + decode_slice], self.spec["state_len"], self.spec["blank_score"])[:3]
                    for i in range(0, ntc.shape[0], decode_slice)]
            moves, seq, qual = (np.concatenate(p) for p in zip(*outs))
        t2 = time.perf_counter()
        return moves, seq, qual, t1 - t0, t2 - t1
