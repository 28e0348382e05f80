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
