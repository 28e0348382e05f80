"""The C-ABI library loads and exports every symbol include/bonito_b200.h declares (no GPU calls)."""
import ctypes
import os
import re

import pytest

from bonito_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "bonito_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared() == sorted(native.SIGNATURES)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(native.lib_path()):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(native.lib_path())
    for name in _declared():
        assert hasattr(lib, name), name
    assert native.version() >= 100
    assert native.load().b200_last_error() == b""
    assert native.lstm_cluster_size(384) == 8 and native.lstm_cluster_size(96) == 1 and native.lstm_cluster_size(100) == 0
    assert native.crf_decode_workspace_bytes(2, 10, 3) > 2 * 11 * 64 * 4


def test_native_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.NativeError, match="CUDA device"):
        native.require()


def test_pipelined_scoring_refuses_cpu_models():
    """No CPU fallback anywhere on the product path: the scoring loop behind basecall() raises on a CPU model."""
    import pytest
    import torch
    from bonito_b200.crf.basecall import score_batches

    class Dummy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))
            self.stride = 6

    with pytest.raises(RuntimeError, match="CUDA"):
        list(score_batches(Dummy(), iter([(0, torch.zeros(2, 1, 60))])))
