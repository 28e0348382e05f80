"""
CTC-CRF model package (`Model`, `basecall` are what `load_symbol` looks up:
`/root/reference/bonito/util.py:223-234`, `/root/reference/bonito/cli/basecaller.py:71`).

Inference-only mirror of `/root/reference/bonito/crf/model.py`: the state graph (`CTC_CRF`),
`SeqdistModel` and `Model` with the `use_koi` swap-in hook.  Training losses (koi.ctc) are out
of scope.
"""

import numpy as np
import torch

from bonito_b200.nn import (Module, Convolution, LinearCRFEncoder, Serial, Permute, layers, to_dict, from_dict,
                            register)


def get_stride(m, stride=1):
    """Total down-sampling of a module tree (reference: bonito/crf/model.py:15-27)."""
    if hasattr(m, "output_stride"):
        return m.output_stride(stride)
    if hasattr(m, "stride"):
        s = m.stride
        if isinstance(s, tuple):
            assert len(s) == 1
            s = s[0]
        return stride * s
    for child in m.children():
        stride = get_stride(child, stride)
    return stride


class CTC_CRF:
    """
    State graph of the k-mer CRF: 4**state_len states, 5 in-edges per state
    (`idx[s, 0] = s` stay; `idx[s, 1+j] = j * 4**(state_len-1) + s // 4` move),
    flat score index `s * 5 + e` (reference: bonito/crf/model.py:30-45).
    """

    def __init__(self, state_len, alphabet):
        self.alphabet = alphabet
        self.state_len = state_len
        self.n_base = len(alphabet[1:])
        n_states = self.n_base ** state_len
        states = torch.arange(n_states)
        moves = states.repeat_interleave(self.n_base).reshape(self.n_base, -1).T
        self.idx = torch.cat([states[:, None], moves], dim=1).to(torch.int32)

    def n_score(self):
        return len(self.alphabet) * self.n_base ** self.state_len

    def reverse_complement(self, scores):
        """Scores of the reverse-complement strand (reference: bonito/crf/model.py:84-96); [T, N, C] layout."""
        T, N, C = scores.shape
        k, nb = self.state_len, self.n_base
        x = scores.reshape(T, N, *([nb] * k), nb + 1)
        blanks = torch.flip(x[..., 0].permute(0, 1, *range(k + 1, 1, -1)).reshape(T, N, -1, 1), [0, 2])
        emissions = torch.flip(
            x[..., 1:].permute(0, 1, *range(k, 1, -1), k + 2, k + 1).reshape(T, N, -1, nb), [0, 2, 3])
        return torch.cat([blanks, emissions], dim=-1).reshape(T, N, -1)

    def path_to_str(self, path):
        letters = np.frombuffer("".join(self.alphabet).encode(), dtype="u1")
        return letters[path[path != 0]].tobytes().decode()


def conv(c_in, c_out, ks, stride=1, bias=False, activation=None, norm=None):
    return Convolution(c_in, c_out, ks, stride=stride, padding=ks // 2, bias=bias, activation=activation, norm=norm)


def rnn_encoder(n_base, state_len, insize=1, first_conv_size=4, stride=5, winlen=19, activation="swish",
                rnn_type="lstm", features=768, scale=5.0, blank_score=None, expand_blanks=True, num_layers=5,
                norm=None):
    """Old-style ([encoder] without `type`) config -> module tree (reference: bonito/crf/model.py:150-162)."""
    rnn = layers[rnn_type]
    return Serial([
        conv(insize, first_conv_size, ks=5, bias=True, activation=activation, norm=norm),
        conv(first_conv_size, 16, ks=5, bias=True, activation=activation, norm=norm),
        conv(16, features, ks=winlen, stride=stride, bias=True, activation=activation, norm=norm),
        Permute([2, 0, 1]),
        *(rnn(features, features, reverse=(num_layers - i) % 2) for i in range(num_layers)),
        LinearCRFEncoder(features, n_base, state_len, activation="tanh", scale=scale,
                         blank_score=blank_score, expand_blanks=expand_blanks),
    ])


@register
class SeqdistModel(Module):
    def __init__(self, encoder, seqdist, n_pre_post_context_bases=None, target_projection=None):
        super().__init__()
        self.seqdist = seqdist
        self.encoder = encoder
        self.stride = get_stride(encoder)
        self.alphabet = seqdist.alphabet
        if n_pre_post_context_bases is None:
            self.n_pre_context_bases, self.n_post_context_bases = seqdist.state_len - 1, 1
        else:
            self.n_pre_context_bases, self.n_post_context_bases = n_pre_post_context_bases
        if target_projection is None:
            self.target_projection = None
        else:
            self.register_buffer("target_projection", torch.tensor([0] + target_projection), persistent=False)
        self._native = None        # set by use_koi(): dict of basecaller settings
        self._plan = None          # built lazily, after the weights are loaded / fused

    @classmethod
    def from_dict(cls, model_dict, layer_types=None):
        kwargs = dict(model_dict, encoder=from_dict(model_dict["encoder"], layer_types),
                      seqdist=CTC_CRF(**model_dict["seqdist"]))
        return cls(**kwargs)

    # -- forward -----------------------------------------------------------------------------------
    def forward(self, x, *args, slot=0):
        """
        Plain module tree ([T, N, C+blanks], the reference's non-koi path) unless `use_koi` armed the
        native engine, in which case the result is [N, T, C] fp16 without blank column and any failure to
        reach the sm_100a kernels raises (no CPU fallback).
        """
        if self._native is None:
            if x.is_cuda:
                # the only CUDA path of this package is the native engine; an eager-torch forward here would be a silent
                # non-native result with a different layout ([T, N, C+blanks])
                raise RuntimeError("bonito_b200: CUDA model called without use_koi(); call model.use_koi(...) "
                                   "(load_model(..., use_koi=True)) or run the module tree on the CPU")
            return self.encoder(x)
        # `slot`: independent buffer set of the native plan, for callers that keep several batches in flight on
        # different streams (score_batches)
        return self.native_plan(x.device if x.is_cuda else None).forward(x, slot=slot)

    def native_plan(self, device=None):
        from bonito_b200 import native
        from bonito_b200.engine import compile_lstm_crf
        from bonito_b200.engine_tf import compile_transformer, find_transformer_encoder
        native.require()
        if device is None:
            device = next(self.parameters()).device
        if torch.device(device).type != "cuda":
            raise native.NativeError("the native path was requested (use_koi) but the model is not on a CUDA device")
        if self._plan is None or self._plan.device != torch.device(device):
            if find_transformer_encoder(self.encoder) is not None:
                self._plan = compile_transformer(self.encoder, device)
            else:
                self._plan = compile_lstm_crf(self.encoder, device, quantize=bool((self._native or {}).get("quantize")))
        return self._plan

    def invalidate_plan(self):
        self._plan = None

    def _apply(self, fn, *args, **kwargs):
        self._plan = None  # .half()/.to() change the tensors the plan was packed from
        return super()._apply(fn, *args, **kwargs)

    def apply(self, fn):
        self._plan = None  # e.g. model.apply(fuse_bn_) rewrites the conv weights
        return super().apply(fn)

    def load_state_dict(self, *args, **kwargs):
        self._plan = None  # the plan holds packed copies of the weights
        return super().load_state_dict(*args, **kwargs)

    def use_koi(self, **kwargs):
        """Arm the B200 engine (the hook `_load_model` calls: bonito/util.py:292-296)."""
        self._native = dict(kwargs)
        self._plan = None

    # -- decode ------------------------------------------------------------------------------------
    def decode_batch(self, x):
        """
        x: scores.  Native layout [N, T, C] (no blanks) on CUDA -> list of N strings, via the sm_100a
        posterior-Viterbi kernel (same maths as the reference's decode_batch, bonito/crf/model.py:196-199).
        """
        from bonito_b200.decode import beam_search, to_str
        if not x.is_cuda:
            raise NotImplementedError("decode_batch runs on the native CUDA decoder only")
        seq, _, _ = beam_search(x.contiguous(), blank_score=self._blank_score())
        return [to_str(row) for row in seq]

    def decode(self, x):
        return self.decode_batch(x.unsqueeze(0))[0]

    def _blank_score(self):
        for m in self.encoder.modules():
            if isinstance(m, LinearCRFEncoder) and m.blank_score is not None:
                return float(m.blank_score)
        return 2.0

    def to_dict(self, include_weights=False):
        if include_weights:
            raise NotImplementedError
        out = {
            "encoder": to_dict(self.encoder),
            "seqdist": {"state_len": self.seqdist.state_len, "alphabet": self.seqdist.alphabet},
            "n_pre_post_context_bases": (self.n_pre_context_bases, self.n_post_context_bases),
        }
        if self.target_projection is not None:
            out["target_projection"] = self.target_projection.tolist()[1:]
        return out


class Model(SeqdistModel):
    """`Model(config)` for `package = "bonito.crf"` configs (reference: bonito/crf/model.py:225-246)."""

    def __init__(self, config):
        seqdist = CTC_CRF(state_len=config["global_norm"]["state_len"], alphabet=config["labels"]["labels"])
        if "type" in config["encoder"]:
            encoder = from_dict(config["encoder"])
        else:
            encoder = rnn_encoder(seqdist.n_base, seqdist.state_len, insize=config["input"]["features"],
                                  **config["encoder"])
        super().__init__(encoder, seqdist, n_pre_post_context_bases=config["input"].get("n_pre_post_context_bases"))
        self.config = config
