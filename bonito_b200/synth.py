"""
Synthetic LSTM-CRF / transformer models and signals (a utility of the package: no kernel or oracle code in here).

There is no network in the build environment, so the real `dna_r10.4.1_e8.2_400bps_{fast,hac}@v5.0.0`
checkpoints cannot be fetched (`/root/reference/bonito/cli/download.py:31-83`); every test and benchmark uses
seeded random weights of the same architecture (shapes: SURVEY.md Appendix A), stored in the reference's own
on-disk format (`config.toml` + `weights_1.tar`) so the same files drive the reference modules, the oracle
and the B200 engine.
"""

import os

import numpy as np
import torch

SHAPES = {
    # name: (hidden, state_len)
    "fast": (96, 3),
    "hac": (384, 4),
    "tiny": (96, 3),
}


def model_spec(name="hac", n_lstm=5, stride=6, winlen=19):
    hidden, state_len = SHAPES[name]
    return dict(
        name=name, hidden=hidden, state_len=state_len, n_lstm=n_lstm,
        convs=[(1, 16, 5, 1, 2, "swish"), (16, 16, 5, 1, 2, "swish"), (16, hidden, winlen, stride, winlen // 2, "tanh")],
        reverse=[bool((i + 1) % 2) for i in range(n_lstm)],  # 1,0,1,0,1 as in @v4.3.toml:62-95
        blank_score=2.0, clamp=(-5.0, 5.0), stride=stride,
    )


def model_config(spec, batchnorm=False, batchsize=32, chunksize=3996, overlap=492):
    """TOML-equivalent dict for `Model(config)` (layout of dna_r10.4.1@v4.3.toml)."""
    sub = []
    for cin, cout, k, s, p, act in spec["convs"]:
        layer = dict(type="convolution", insize=cin, size=cout, bias=True, winlen=k, stride=s, padding=p, activation=act)
        if batchnorm:
            layer["norm"] = "batchnorm"
        sub.append(layer)
    sub.append(dict(type="permute", dims=[2, 0, 1]))
    for i in range(spec["n_lstm"]):
        sub.append(dict(type="lstm", size=spec["hidden"], insize=spec["hidden"], bias=True, reverse=int(spec["reverse"][i])))
    crf = dict(type="linearcrfencoder", insize=spec["hidden"], n_base=4, state_len=spec["state_len"], bias=False,
               blank_score=spec["blank_score"])
    if spec.get("crf_activation") is not None:          # old-style head: tanh + scale instead of a Clamp layer
        crf["activation"] = spec["crf_activation"]
    if spec.get("crf_scale") is not None:
        crf["scale"] = spec["crf_scale"]
    sub.append(crf)
    if spec.get("clamp") is not None:
        sub.append(dict(type="clamp", min=spec["clamp"][0], max=spec["clamp"][1]))
    return {
        "model": {"package": "bonito.crf"},
        "labels": {"labels": ["N", "A", "C", "G", "T"]},
        "input": {"features": 1},
        "global_norm": {"state_len": spec["state_len"]},
        "qscore": {"scale": 1.05, "bias": 0.2},
        # picoampere input, standardised with fixed statistics (v4.3+/v5 LSTM configs)
        "scaling": {"strategy": "pa"},
        "standardisation": {"standardise": 1, "mean": 93.7, "stdev": 23.5},
        "encoder": {"type": "serial", "sublayers": sub},
        "basecaller": {"batchsize": batchsize, "chunksize": chunksize, "overlap": overlap},
    }


def _orthogonal_blocks(rows, cols, block, gen, gain):
    w = torch.empty(rows, cols)
    for r in range(0, rows, block):
        q, _ = torch.linalg.qr(torch.randn(max(block, cols), max(block, cols), generator=gen))
        w[r:r + block] = q[:block, :cols]
    return w * gain


def make_weights(spec, seed=25, conv_gain=2.5, lstm_gain=1.5, head_gain=6.0, fp16_values=True):
    """
    Seeded, non-degenerate weights (oracle naming).  The reference's own init (orthogonal LSTM blocks,
    0.5*truncated-normal input bias, zero state bias: bonito/nn.py:362-390) with gains chosen so that
    decoded sequences vary from chunk to chunk, the +-5 clamp rarely saturates (SURVEY.md hard part H5) and the
    recurrence stays well conditioned (an LSTM gain of 3 makes the stack chaotic: a 1e-3 input perturbation grows to
    O(1) score differences, so no two half-precision implementations could agree; at 1.5 it shrinks).
    With `fp16_values` every tensor is rounded to fp16 (what `model.half()` feeds every implementation).
    """
    gen = torch.Generator().manual_seed(seed)
    H = spec["hidden"]
    w = {}
    for i, (cin, cout, k, _, _, _) in enumerate(spec["convs"]):
        fan_in = cin * k
        w[f"conv{i}.weight"] = torch.randn(cout, cin, k, generator=gen) * (conv_gain / fan_in ** 0.5)
        w[f"conv{i}.bias"] = torch.randn(cout, generator=gen) * 0.1
    for i in range(spec["n_lstm"]):
        w[f"lstm{i}.w_ih"] = _orthogonal_blocks(4 * H, H, H, gen, lstm_gain)
        w[f"lstm{i}.w_hh"] = _orthogonal_blocks(4 * H, H, H, gen, lstm_gain)
        w[f"lstm{i}.b_ih"] = 0.5 * torch.randn(4 * H, generator=gen).clamp(-2, 2)
        w[f"lstm{i}.b_hh"] = torch.zeros(4 * H)
    C = 4 ** (spec["state_len"] + 1)
    w["crf.weight"] = torch.randn(C, H, generator=gen) * (head_gain / H ** 0.5)
    if fp16_values:
        w = {k: v.half().float() for k, v in w.items()}
    return w


def state_dict_from_weights(spec, weights, prefix="encoder."):
    """Oracle naming -> the module tree's state_dict keys (SURVEY.md Appendix A 'State-dict names')."""
    sd = {}
    n_conv = len(spec["convs"])
    for i in range(n_conv):
        sd[f"{prefix}{i}.conv.weight"] = weights[f"conv{i}.weight"]
        sd[f"{prefix}{i}.conv.bias"] = weights[f"conv{i}.bias"]
    base = n_conv + 1  # + Permute
    for i in range(spec["n_lstm"]):
        sd[f"{prefix}{base + i}.rnn.weight_ih_l0"] = weights[f"lstm{i}.w_ih"]
        sd[f"{prefix}{base + i}.rnn.weight_hh_l0"] = weights[f"lstm{i}.w_hh"]
        sd[f"{prefix}{base + i}.rnn.bias_ih_l0"] = weights[f"lstm{i}.b_ih"]
        sd[f"{prefix}{base + i}.rnn.bias_hh_l0"] = weights[f"lstm{i}.b_hh"]
    sd[f"{prefix}{base + spec['n_lstm']}.linear.weight"] = weights["crf.weight"]
    return sd


def write_model_dir(dirname, spec, weights, **config_kwargs):
    """Write `config.toml` + `weights_1.tar` in the reference's format (bonito/util.py:271-305)."""
    import toml
    os.makedirs(dirname, exist_ok=True)
    with open(os.path.join(dirname, "config.toml"), "w") as fh:
        toml.dump(model_config(spec, **config_kwargs), fh)
    torch.save(state_dict_from_weights(spec, weights), os.path.join(dirname, "weights_1.tar"))
    return dirname


def squiggle(n, length, seed=25, dwell=10.0, noise=0.15):
    """
    Piecewise-constant synthetic nanopore signal, ~N(0,1) after standardisation:
    levels ~ N(0,1) held for geometric dwell times (mean `dwell` samples) plus N(0, noise^2).
    """
    rng = np.random.default_rng(seed)
    out = np.empty((n, length), dtype=np.float32)
    for i in range(n):
        n_levels = int(length / dwell * 2) + 8
        dwells = rng.geometric(1.0 / dwell, size=n_levels)
        levels = rng.standard_normal(n_levels).astype(np.float32)
        sig = np.repeat(levels, dwells)[:length]
        out[i] = sig + noise * rng.standard_normal(length).astype(np.float32)
    return torch.from_numpy(out)[:, None, :]


def gaussian_signal(n, length, seed=25):
    gen = torch.Generator().manual_seed(seed)
    return torch.randn(n, 1, length, generator=gen)


# ---------------------------------------------------------------------------------------------------
# transformer (sup v5.0) shapes: /root/reference/bonito/models/configs/dna_r10.4.1@v5.0.toml
# ---------------------------------------------------------------------------------------------------

def sup_spec(depth=18, d_model=512, nhead=8, dim_feedforward=2048, state_len=5):
    alpha = round((2 * depth) ** 0.25, 7)
    beta = round((8 * depth) ** (-1 / 4), 7)
    return dict(
        name="sup", depth=depth, d_model=d_model, nhead=nhead, dim_feedforward=dim_feedforward, state_len=state_len,
        convs=[(1, 64, 5, 1, 2, "swish"), (64, 64, 5, 1, 2, "swish"), (64, 128, 9, 3, 4, "swish"),
               (128, 128, 9, 2, 4, "swish"), (128, d_model, 5, 2, 2, "swish")],
        alpha=alpha, beta=beta, window=(127, 128), scale=5.0, blank_score=2.0, stride=6,
    )


def sup_config(spec, batchnorm=False, batchsize=32, chunksize=12000, overlap=600):
    convs = []
    for cin, cout, k, s, p, act in spec["convs"]:
        layer = dict(type="convolution", insize=cin, size=cout, bias=True, winlen=k, stride=s, padding=p, activation=act)
        if batchnorm:
            layer["norm"] = "batchnorm"
        convs.append(layer)
    convs.append(dict(type="permute", dims=[0, 2, 1]))
    enc = {
        "type": "namedserial",
        "conv": {"type": "serial", "sublayers": convs},
        "transformer_encoder": {"type": "stack", "depth": spec["depth"], "layer": {
            "type": "transformerencoderlayer", "d_model": spec["d_model"], "nhead": spec["nhead"],
            "dim_feedforward": spec["dim_feedforward"], "deepnorm_alpha": spec["alpha"], "deepnorm_beta": spec["beta"],
            "attn_window": list(spec["window"])}},
        "upsample": {"type": "linearupsample", "d_model": spec["d_model"], "scale_factor": 2},
        "crf": {"type": "linearcrfencoder", "insize": spec["d_model"], "n_base": 4, "state_len": spec["state_len"],
                "bias": False, "scale": spec["scale"], "blank_score": spec["blank_score"], "expand_blanks": True,
                "permute": [1, 0, 2]},
    }
    return {
        "model": {"type": "seqdistmodel", "package": "bonito.transformer",
                  "seqdist": {"state_len": spec["state_len"], "alphabet": ["N", "A", "C", "G", "T"]}, "encoder": enc},
        "qscore": {"scale": 1.05, "bias": 1.3},
        "basecaller": {"batchsize": batchsize, "chunksize": chunksize, "overlap": overlap},
    }


def make_sup_weights(spec, seed=25, conv_gain=1.8, head_gain=0.55, fp16_values=True):
    """Seeded weights with the reference's initialisation scheme (xavier with the DeepNorm beta gain on the value /
    output / feed-forward projections: bonito/transformer/model.py:116-123; RMSNorm weights 1), state-dict names
    relative to `encoder.`."""
    gen = torch.Generator().manual_seed(seed)
    d, ff, beta = spec["d_model"], spec["dim_feedforward"], spec["beta"]
    w = {}
    for i, (cin, cout, k, _, _, _) in enumerate(spec["convs"]):
        w[f"conv.{i}.conv.weight"] = torch.randn(cout, cin, k, generator=gen) * (conv_gain / (cin * k) ** 0.5)
        w[f"conv.{i}.conv.bias"] = torch.randn(cout, generator=gen) * 0.1

    def xavier(rows, cols, gain):
        return torch.randn(rows, cols, generator=gen) * gain * (2.0 / (rows + cols)) ** 0.5

    for l in range(spec["depth"]):
        p = f"transformer_encoder.{l}."
        w[p + "self_attn.Wqkv.weight"] = torch.cat([xavier(2 * d, d, 1.0) * 3.0, xavier(d, d, beta)])
        w[p + "self_attn.out_proj.weight"] = xavier(d, d, beta)
        w[p + "self_attn.out_proj.bias"] = torch.randn(d, generator=gen) * 0.02
        w[p + "ff.fc1.weight"] = xavier(2 * ff, d, beta)
        w[p + "ff.fc2.weight"] = xavier(d, ff, beta)
        w[p + "norm1.weight"] = torch.ones(d)
        w[p + "norm2.weight"] = torch.ones(d)
    w["upsample.linear.weight"] = xavier(2 * d, d, 1.0)
    w["upsample.linear.bias"] = torch.randn(2 * d, generator=gen) * 0.02
    C = 4 ** (spec["state_len"] + 1)
    w["crf.linear.weight"] = torch.randn(C, d, generator=gen) * (head_gain / d ** 0.5)
    if fp16_values:
        w = {k: v.half().float() for k, v in w.items()}
    return w


def sup_state_dict(spec, weights, prefix="encoder."):
    sd = {prefix + k: v for k, v in weights.items()}
    for l in range(spec["depth"]):
        sd[f"{prefix}transformer_encoder.{l}.deepnorm_alpha"] = torch.tensor(spec["alpha"])
    return sd
