"""CLI surface (reference: test/test_cli.py answers `-h` for every subcommand), reader and FASTQ writer."""
import io
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_basecaller_help_and_flag_surface():
    out = subprocess.run([sys.executable, "-m", "bonito_b200", "basecaller", "-h"], cwd=ROOT, capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("usage: bonito_b200 basecaller")
    from bonito_b200.cli.basecaller import argparser
    flags = {a for act in argparser()._actions for a in act.option_strings}
    reference_flags = {"--reference", "--read-ids", "--device", "--seed", "--weights", "--skip", "--no-trim", "--save-ctc",
                       "--revcomp", "--rna", "--recursive", "--quantize", "--no-quantize", "--overlap", "--chunksize",
                       "--batchsize", "--max-reads", "--min-qscore", "--min-accuracy-save-ctc", "--alignment-threads",
                       "--mm2-preset", "-v", "--verbose"}     # bonito/cli/basecaller.py:168-199
    assert flags == reference_flags
    args = argparser().parse_args(["model", "reads"])
    assert (args.device, args.seed, args.weights, args.quantize, args.chunksize) == ("cuda", 25, 0, None, None)


def test_reader_npy_trim_and_normalisation(tmp_path):
    from bonito_b200.reader import Read, Reader, normalisation, trim
    rng = np.random.default_rng(0)
    for i in range(3):
        np.save(tmp_path / f"read{i}.npy", (90 + 20 * rng.standard_normal(3000 + 100 * i)).astype(np.float32))
    reader = Reader(str(tmp_path))
    assert reader.fmt == "npy"
    reads = list(reader.get_reads(str(tmp_path), scaling_strategy={"strategy": "pa"},
                                  norm_params={"standardise": 1, "mean": 90.0, "stdev": 20.0}))
    assert [r.read_id for r in reads] == ["read0", "read1", "read2"]
    assert abs(float(reads[0].signal.mean())) < 0.1 and abs(float(reads[0].signal.std()) - 1) < 0.1
    assert reads[0].signal.dtype == np.float32 and len(reads[0].signal) == 3000 - reads[0].trimmed_samples
    only = list(reader.get_reads(str(tmp_path), read_ids={"read1"}, scaling_strategy={"strategy": "pa"},
                                 norm_params={"standardise": 0}))
    assert [r.read_id for r in only] == ["read1"]
    skipped = list(reader.get_reads(str(tmp_path), read_ids={"read1"}, skip=True))
    assert [r.read_id for r in skipped] == ["read0", "read2"]
    # quantile scaling defaults (bonito/reader.py:17-20,157-163)
    sig = np.linspace(0, 1000, 1001)
    shift, scale = normalisation(sig)
    assert abs(shift - 0.51 * (200 + 900)) < 1e-6 and abs(scale - 0.53 * 700) < 1e-6
    with pytest.raises(ValueError):
        normalisation(sig, {"strategy": "pa"}, None)
    # trim: a stall above threshold followed by quiet signal is cut at the end of the first quiet window
    s = np.concatenate([np.full(210, 5.0), np.zeros(2000)]).astype(np.float32)
    assert trim(s) == 250 and trim(np.zeros(1000, dtype=np.float32)) == 10
    with pytest.raises(FileNotFoundError):
        Reader(str(tmp_path / "nothing"))


def test_fastq_writer_filters_and_logs():
    from bonito_b200.io import Writer

    class R:
        def __init__(self, rid, n):
            self.read_id, self.signal, self.trimmed_samples = rid, np.zeros(n, dtype=np.float32), 5

    results = [(R("a", 100), {"sequence": "ACGT", "qstring": "5555"}), (R("b", 50), {"sequence": "", "qstring": ""}),
               (R("c", 70), {"sequence": "GG", "qstring": "##"})]
    fd = io.StringIO()
    w = Writer(iter(results), fd=fd, min_qscore=10)
    w.start(); w.join()
    assert w.error is None
    assert fd.getvalue() == "@a\nACGT\n+\n5555\n"
    assert w.log == [("a", 105)]


def test_sam_writer_record_layout_and_move_table():
    """Unaligned SAM text as the reference lays it out (bonito/io.py:136-166,441-456; documentation/SAM.md): flag 4, NM:i:0,
    RG / qs / ns / ts, the read's tag data and the move table mv:B:c,<stride>,<moves>."""
    from bonito_b200.io import Writer, encode_moves, sam_record
    from bonito_b200.reader import Read
    assert encode_moves(np.array([0, 1, 0, 1, 1], dtype=np.int8), 5) == "5,0,1,0,1,1"      # the reference's doctest
    assert sam_record("r", "ACGT", "5555", None, tags=["qs:i:20"]) == "r\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\t5555\tNM:i:0\tqs:i:20"
    read = Read("read7", 100.0 + np.arange(300, dtype=np.float32) % 7, filename="read7.npy", do_trim=False,
                scaling_strategy={"strategy": "pa"}, norm_params={"standardise": 1, "mean": 100.0, "stdev": 2.0},
                meta={"run_id": "runA", "channel": 12, "mux": 3, "read_number": 9})
    res = {"sequence": "ACGTA", "qstring": "55555", "moves": np.array([1, 0, 1, 1, 0, 1, 0, 1], dtype=np.int8), "stride": 6}
    fd = io.StringIO()
    w = Writer(iter([(read, res)]), fd=fd, mode="w", groups=[read.readgroup("model_x")], group_key="model_x")
    w.start(); w.join()
    assert w.error is None
    lines = fd.getvalue().rstrip("\n").split("\n")
    assert lines[0].startswith("@HD\tVN:1.5\tSO:unknown\tob:0.0.2") and lines[1].startswith("@PG\tID:basecaller")
    assert lines[2].startswith("@RG\tID:runA_model_x\tPL:ONT")
    f = lines[3].split("\t")
    assert f[:12] == ["read7", "4", "*", "0", "0", "*", "*", "0", "0", "ACGTA", "55555", "NM:i:0"]
    tags = f[12:]
    assert tags[:4] == ["RG:Z:runA_model_x", "qs:i:20", "ns:i:300", "ts:i:0"]
    assert "mx:i:3" in tags and "ch:i:12" in tags and "rn:i:9" in tags and "f5:Z:read7.npy" in tags and "sv:Z:pa" in tags
    assert tags[-1] == "mv:B:c,6,1,0,1,1,0,1,0,1"
    with pytest.raises(ValueError):
        Writer(iter([]), mode="wb")
