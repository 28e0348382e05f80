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


def test_chunk_count_is_the_host_functions_chunk_count():
    """b200_chunk_count (host arithmetic of the device-side chunk()) against bonito_b200.util.chunk, whose windows are pinned
    by the reference-generated fixture (tests/golden/host_logic.npz)."""
    import torch
    from bonito_b200.util import chunk
    lib = native.load()
    for length in (1, 37, 499, 500, 3999, 4000, 4001, 7500, 11500, 12000, 40000, 123457):
        for chunksize, overlap in ((4000, 500), (3996, 498), (9996, 498), (1000, 0), (600, 599)):
            assert lib.b200_chunk_count(length, chunksize, overlap) == chunk(torch.zeros(length), chunksize, overlap).shape[0], \
                (length, chunksize, overlap)
    assert lib.b200_chunk_count(100, 50, 50) == 0 and lib.b200_chunk_count(0, 50, 5) == 0
