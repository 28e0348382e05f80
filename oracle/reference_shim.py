"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Import the *real* reference modules from /root/reference.

Only usable in the authoring container (the GPU box has no /root/reference): used by
`oracle/make_golden.py` to generate `tests/golden/*.npz` and by tests marked `needs_reference`.
`import bonito` itself fails there (mappy / pysam / koi are absent), so a namespace package pointing at the
reference tree is installed together with minimal stand-ins for the absent third-party modules.  The
stand-in for `koi.ctc.SequenceDist.posteriors` is the oracle's restatement (it cannot be the original:
ont-koi is a closed binary wheel) -- everything else executed is the reference's own code.
"""

import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "bonito"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def load():
    """Return a namespace with the reference's nn / util / crf.model modules."""
    if "bonito" in sys.modules and getattr(sys.modules["bonito"], "__shim__", False):
        pkg = sys.modules["bonito"]
        return types.SimpleNamespace(nn=sys.modules["bonito.nn"], util=sys.modules["bonito.util"],
                                     crf_model=sys.modules["bonito.crf.model"])
    if not available():
        raise RuntimeError("reference tree not present")
    from oracle import crf_oracle

    class _Semiring:
        def __init__(self, name, one):
            self.name, self.one = name, one

    Log, Max = _Semiring("log", 0.0), _Semiring("max", 0.0)

    class SequenceDist:
        """Stand-in for koi.ctc.SequenceDist: only `posteriors` (restated by the oracle)."""

        def posteriors(self, scores, S=Log):
            T, N, C = scores.shape
            Ms = scores.detach().double().numpy().reshape(T, N, -1, self.n_base + 1)
            idx = self.idx.numpy().astype(np.int64)
            if S is Log:
                post = crf_oracle.posteriors(Ms, idx)
            else:
                states, edges = crf_oracle.viterbi_edges(Ms, idx)
                post = np.zeros_like(Ms)
                t, n = np.meshgrid(np.arange(T), np.arange(N), indexing="ij")
                post[t, n, states, edges] = 1.0
            return torch.from_numpy(post.reshape(T, N, C)).to(scores.dtype)

    _stub("koi")
    _stub("koi.lstm", update_graph=lambda enc, **kw: enc)
    _stub("koi.ctc", SequenceDist=SequenceDist, Max=Max, Log=Log, semiring=_Semiring, logZ_cu=None,
          viterbi_alignments=None, logZ_cu_sparse=None, bwd_scores_cu_sparse=None, fwd_scores_cu_sparse=None)
    _stub("koi.decode", beam_search=None, to_str=None)
    _stub("parasail")
    pkg = types.ModuleType("bonito")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "bonito")]
    pkg.__shim__ = True
    sys.modules["bonito"] = pkg
    import importlib
    nn = importlib.import_module("bonito.nn")
    util = importlib.import_module("bonito.util")
    crf_model = importlib.import_module("bonito.crf.model")
    return types.SimpleNamespace(nn=nn, util=util, crf_model=crf_model)


def load_transformer():
    """
    The reference's transformer package, runnable on the CPU: its module classes are used unchanged; the pieces that
    only exist as Triton / CUDA kernels are routed to flash-attn's own torch reference functions:
      * `attn_func` takes its SDPA + sliding_window_mask branch (bonito/transformer/model.py:61-65)
        (the capability probe at :59 is answered with (7, 0));
      * RotaryEmbedding.forward -> flash_attn.layers.rotary.apply_rotary_emb_torch with the module's cos/sin cache
        recipe (fp32 tables cast to the qkv dtype);
      * RMSNorm.forward -> flash_attn.ops.triton.layer_norm.rms_norm_ref(upcast=True);
      * GatedMlp.forward -> fc1, chunk, swiglu_fwd formula (flash_attn/ops/activations.py:107-111), fc2.
    """
    load()
    import importlib
    # flash-attn's Triton modules probe the CUDA device at import time; answer the probes so they import on a CPU box
    if not torch.cuda.is_available():
        torch.cuda.current_device = lambda: 0
        torch.cuda.get_device_properties = lambda *a, **k: types.SimpleNamespace(
            warp_size=32, multi_processor_count=1, major=7, minor=0, name="cpu-shim")
    from flash_attn.layers.rotary import apply_rotary_emb_torch
    from flash_attn.ops.triton.layer_norm import rms_norm_ref
    torch.cuda.get_device_capability = lambda *a, **k: (7, 0)
    tm = importlib.import_module("bonito.transformer.model")

    def rotary_forward(self, qkv, seqlen_offset=0, max_seqlen=None):
        T = qkv.shape[1]
        inv_freq = 1.0 / (self.base ** (torch.arange(0, self.dim, 2, dtype=torch.float32) / self.dim))
        freqs = torch.outer(torch.arange(T, dtype=torch.float32), inv_freq)
        cos, sin = torch.cos(freqs).to(qkv.dtype), torch.sin(freqs).to(qkv.dtype)
        q = apply_rotary_emb_torch(qkv[:, :, 0], cos, sin)
        k = apply_rotary_emb_torch(qkv[:, :, 1], cos, sin)
        return torch.stack([q, k, qkv[:, :, 2]], dim=2)

    def rmsnorm_forward(self, x, residual=None, prenorm=False, residual_in_fp32=False):
        return rms_norm_ref(x, self.weight, None, residual=residual, eps=self.eps, upcast=True)

    def gated_forward(self, x):
        y, gate = self.fc1(x).chunk(2, dim=-1)
        g = gate.float()
        return self.fc2((g * y.float() / (1.0 + torch.exp(-g))).to(x.dtype))

    tm.RotaryEmbedding.forward = rotary_forward
    tm.RMSNorm.forward = rmsnorm_forward
    tm.GatedMlp.forward = gated_forward
    return tm
