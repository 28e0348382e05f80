"""CPU tests of the host-side mirror (bonito_b200.util / nn / crf.model) against fixtures generated from the
reference's own functions (oracle/make_golden.py) and the known answers of SURVEY.md Appendix A."""
import json
import os

import numpy as np
import pytest
import torch

from bonito_b200 import nn as bnn
from bonito_b200 import util
from bonito_b200.crf.model import CTC_CRF, Model, get_stride


@pytest.fixture(scope="module")
def host(golden_dir):
    return np.load(os.path.join(golden_dir, "host_logic.npz"))


def test_chunk_and_stitch_match_reference(host):
    for i, (L, cs, ov, stride) in enumerate(host["cases"]):
        L, cs, ov, stride = int(L), int(cs), int(ov), int(stride)
        sig = torch.arange(L, dtype=torch.float32)
        chunks = util.chunk(sig, cs, ov)
        assert list(chunks.shape) == list(host[f"chunk_shape_{i}"])
        np.testing.assert_array_equal(chunks[:, 0, 0].numpy(), host[f"chunk_first_{i}"])
        frames = chunks[:, 0, ::stride]
        if L < cs:
            from bonito_b200.crf.basecall import stitch_results
            got = stitch_results(frames, L, cs, ov, stride)
        else:
            got = util.stitch(frames, cs, ov, L, stride)
            np.testing.assert_array_equal(util.stitch(frames, cs, ov, L, stride, reverse=True).numpy(),
                                          host[f"stitch_rev_{i}"])
        np.testing.assert_array_equal(got.numpy(), host[f"stitch_{i}"])


def test_known_answers_from_survey():
    # (L, chunk, overlap, stride) -> (n_chunks, T, stitched_len)
    for (L, cs, ov, st), (n, T, out_len) in {
        (25000, 3996, 492, 6): (7, 666, 4166),
        (100000, 9996, 492, 6): (11, 1666, 16666),
        (60000, 12000, 600, 6): (6, 2000, 10000),
    }.items():
        chunks = util.chunk(torch.arange(L, dtype=torch.float32), cs, ov)
        frames = chunks[:, 0, ::st]
        assert chunks.shape[0] == n and frames.shape[1] == T
        assert util.stitch(frames, cs, ov, L, st).shape[0] == out_len


def test_chunk_edge_cases():
    sig = torch.arange(10, dtype=torch.float32)
    assert util.chunk(sig, 0, 0).shape == (1, 1, 10)            # chunksize 0: whole read
    tiled = util.chunk(sig, 25, 5)                              # short read is tiled
    assert tiled.shape == (1, 1, 25)
    np.testing.assert_array_equal(tiled[0, 0].numpy(), np.concatenate([np.arange(10)] * 2 + [np.arange(5)]))
    exact = util.chunk(torch.arange(40, dtype=torch.float32), 20, 10)  # windows tile exactly: no stub chunk
    assert exact.shape == (3, 1, 20)


def test_batchify_unbatchify(host):
    items = [(f"r{j}", torch.arange(n * 3, dtype=torch.float32).reshape(n, 3) + 100 * j) for j, n in enumerate([3, 5, 1, 7])]
    batches = list(util.batchify(items, 4))
    keys = [[[k, list(se)] for k, se in ks] for ks, _ in batches]
    assert keys == json.loads(str(host["batchify_keys"]))
    assert [v.shape[0] for _, v in batches] == list(host["batchify_sizes"])
    rebuilt = list(util.unbatchify(batches))
    assert [k for k, _ in rebuilt] == json.loads(str(host["unbatchify_keys"]))
    np.testing.assert_array_equal(torch.cat([v for _, v in rebuilt]).numpy(), host["unbatchify_cat"])
    # dict values (what compute_scores returns) and an empty stream
    d_batches = [((("a", (0, 3)), ("b", (3, 4))), {"m": torch.zeros(4, 2), "s": torch.ones(4, 2)}),
                 ((("b", (0, 1)),), {"m": torch.zeros(1, 2), "s": torch.ones(1, 2)})]
    out = list(util.unbatchify(d_batches))
    assert [k for k, _ in out] == ["a", "b"] and out[0][1]["m"].shape == (3, 2) and out[1][1]["s"].shape == (2, 2)
    assert list(util.batchify([], 4)) == []


def test_crf_state_graph(host):
    for k in (3, 4, 5):
        np.testing.assert_array_equal(CTC_CRF(k, ["N", "A", "C", "G", "T"]).idx.numpy(), host[f"idx_k{k}"])
        S = 4 ** k
        s = np.arange(S)
        closed = np.stack([s] + [j * (S // 4) + s // 4 for j in range(4)], axis=1)
        np.testing.assert_array_equal(host[f"idx_k{k}"], closed)


def test_registry_names_and_round_trip():
    import bonito_b200.crf  # noqa: F401  (registers seqdistmodel)
    expected = {"relu", "tanh", "linear", "swish", "clamp", "serial", "stack", "namedserial", "linearupsample",
                "reverse", "batchnorm", "convolution", "linearcrfencoder", "permute", "lstm", "seqdistmodel"}
    assert expected <= set(bnn.layers)
    from oracle import synth
    cfg = synth.model_config(synth.model_spec("fast", n_lstm=2), batchnorm=True)
    enc = bnn.from_dict(cfg["encoder"])
    descr = bnn.to_dict(enc)
    # to_dict spells out LinearCRFEncoder defaults (scale / expand_blanks), exactly as the reference does
    assert descr["sublayers"][:-2] == cfg["encoder"]["sublayers"][:-2]
    assert descr["sublayers"][-2] == {**cfg["encoder"]["sublayers"][-2], "scale": None, "expand_blanks": True}
    assert bnn.to_dict(bnn.from_dict(descr)) == descr
    assert get_stride(enc) == 6
    with pytest.raises(Exception, match="Failed to build layer"):
        bnn.from_dict({"type": "convolution", "insize": 1})
    # sublayers given as a single dict, and pass-through of built objects
    rev = bnn.from_dict({"type": "reverse", "sublayers": {"type": "lstm", "size": 8, "insize": 8}})
    assert isinstance(rev.layer, bnn.LSTM)
    assert bnn.from_dict(rev) is rev


def test_register_third_party_layer():
    @bnn.register
    class Doubler(torch.nn.Module):
        def forward(self, x):
            return 2 * x
    assert bnn.layers["doubler"] is Doubler and Doubler.name == "doubler"
    layer = bnn.from_dict({"type": "serial", "sublayers": [{"type": "doubler"}]})
    assert float(layer(torch.ones(1))) == 2.0
    del bnn.layers["doubler"]


def test_load_model_from_reference_format(tmp_path):
    from oracle import synth
    spec = synth.model_spec("fast", n_lstm=2)
    weights = synth.make_weights(spec, seed=3)
    d = synth.write_model_dir(str(tmp_path / "m"), spec, weights, chunksize=4000, overlap=500, batchsize=8)
    model = util.load_model(d, "cpu", half=False, use_koi=False)
    assert isinstance(model, Model) and model.stride == 6 and model.alphabet == ["N", "A", "C", "G", "T"]
    sd = model.state_dict()
    assert torch.equal(sd["encoder.4.rnn.weight_hh_l0"], weights["lstm0.w_hh"])
    # use_koi trims chunksize to a stride multiple and overlap to an even multiple (bonito/util.py:288-291)
    model = util.load_model(d, "cpu", half=True, use_koi=True)
    assert model.config["basecaller"]["chunksize"] == 3996 and model.config["basecaller"]["overlap"] == 492
    assert model._native == {"batchsize": 8, "chunksize": 3996, "quantize": False}
    with pytest.raises(Exception):  # native path armed + no CUDA device => loud failure, never a CPU fallback
        model(torch.zeros(1, 1, 3996, dtype=torch.float16))
    # CLI precedence: flag > [basecaller] > defaults
    cfg = util.set_config_defaults({"basecaller": {"chunksize": 10000}}, chunksize=None, batchsize=16, overlap=None)
    assert cfg["basecaller"] == {"chunksize": 10000, "overlap": 500, "batchsize": 16, "quantize": False}
    with pytest.raises(FileNotFoundError):
        util.get_last_checkpoint(str(tmp_path))


def test_match_names_by_shape():
    from oracle import synth
    spec = synth.model_spec("fast", n_lstm=1)
    model = Model(synth.model_config(spec))
    renamed = {f"module.x{i}": v for i, (k, v) in enumerate(model.state_dict().items())}
    mapping = util.match_names(renamed, model)
    assert list(mapping.values()) == list(model.state_dict().keys())


def test_phred_and_mean_qscore():
    assert util.phred(0.9) == chr(10 + 33) and util.phred(1.0) == chr(40 + 33)
    assert util.phred(0.99, scale=1.05, bias=0.2) == chr(int(np.round(20 * 1.05 + 0.2)) + 33)
    assert abs(util.mean_qscore_from_qstring("5" * 10) - 20.0) < 1e-9


def test_reverse_complement_matches_reference(golden_dir):
    """CTC_CRF.reverse_complement against the reference's own function (tests/golden/revcomp.npz), k = 3, 4, 5; and the
    definition's involution property."""
    from bonito_b200.crf.model import CTC_CRF
    gold = np.load(os.path.join(golden_dir, "revcomp.npz"))
    for k in (3, 4, 5):
        sd = CTC_CRF(k, ["N", "A", "C", "G", "T"])
        x = torch.from_numpy(gold[f"in_k{k}"])
        got = sd.reverse_complement(x)
        assert torch.equal(got, torch.from_numpy(gold[f"out_k{k}"]))
        assert torch.equal(sd.reverse_complement(got), x)


def test_fmt_rna_flips_sequence_and_qstring():
    """fmt(..., rna=True) reverses sequence and qstring (bonito/crf/basecall.py:48-55); moves stay as they are."""
    from bonito_b200.crf.basecall import fmt
    seq = torch.tensor([0, 65, 0, 67, 71, 0, 84], dtype=torch.uint8)
    qs = torch.tensor([0, 40, 0, 41, 42, 0, 43], dtype=torch.uint8)
    mv = torch.tensor([0, 1, 0, 1, 1, 0, 1], dtype=torch.uint8)
    attrs = {"sequence": seq, "qstring": qs, "moves": mv}
    dna, rna = fmt(6, attrs), fmt(6, attrs, rna=True)
    assert dna["sequence"] == "ACGT" and dna["qstring"] == "()*+"
    assert rna["sequence"] == "TGCA" and rna["qstring"] == "+*)(" and rna["stride"] == 6
    assert np.array_equal(rna["moves"], mv.numpy())
