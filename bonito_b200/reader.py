"""
Read ingestion for `bonito_b200 basecaller`.

The reference picks a pod5 or fast5 reader by globbing the reads directory (`/root/reference/bonito/reader.py:23-48`);
both need third-party libraries (pod5, ont_fast5_api) that are optional here.  Supported inputs:
  * `*.pod5`  through the `pod5` package when it is importable (signal in pA = scale * (raw + offset));
  * `*.npy`   one read per file: a 1-D array of picoampere samples, read id = file stem (what the tests and the
              synthetic benchmarks use; no third-party dependency).
Trimming and normalisation follow `bonito/reader.py:122-166` (trim of the leading stall, pA standardisation or
quantile scaling).
"""

import os
from glob import glob

import numpy as np

__default_norm_params__ = {"quantile_a": 0.2, "quantile_b": 0.9, "shift_multiplier": 0.51, "scale_multiplier": 0.53}


def trim(signal, window_size=40, threshold=2.4, min_trim=10, min_elements=3, max_samples=8000, max_trim=0.3):
    """Number of leading samples to drop: end of the first run of windows with > min_elements samples above threshold."""
    limit = min(max_samples, len(signal))
    seen_peak = False
    for pos in range(limit // window_size):
        start = pos * window_size + min_trim
        end = start + window_size
        window = signal[start:end]
        if seen_peak or np.count_nonzero(window > threshold) > min_elements:
            seen_peak = True
            if window[-1] > threshold:
                continue
            if end >= limit or end / len(signal) > max_trim:
                return min_trim
            return end
    return min_trim


def normalisation(sig, scaling_strategy=None, norm_params=None):
    """(shift, scale) for `(sig - shift) / scale`: pA standardisation from the config, else quantile scaling."""
    strategy = scaling_strategy.get("strategy") if scaling_strategy else None
    if strategy == "pa":
        if norm_params and norm_params.get("standardise") == 1:
            return norm_params.get("mean"), norm_params.get("stdev")
        if norm_params and norm_params.get("standardise") == 0:
            return 0.0, 1.0
        raise ValueError("Picoampere scaling requested, but standardisation flag not provided")
    if strategy in (None, "quantile"):
        p = norm_params or __default_norm_params__
        qa, qb = np.quantile(sig, [p["quantile_a"], p["quantile_b"]])
        return max(10, p["shift_multiplier"] * (qa + qb)), max(1.0, p["scale_multiplier"] * (qb - qa))
    raise ValueError(f"Scaling strategy {strategy} not supported; choose quantile or pa.")


class Read:
    """What `basecall()` needs (`read_id`, float32 `signal`) plus the bookkeeping the writers use."""

    def __init__(self, read_id, pa_signal, filename="", do_trim=True, scaling_strategy=None, norm_params=None, meta=None):
        meta = meta or {}
        self.read_id, self.filename = str(read_id), filename
        pa = np.asarray(pa_signal, dtype=np.float32)
        self.num_samples = len(pa)
        self.shift, self.scale = normalisation(pa, scaling_strategy, norm_params)
        self.trimmed_samples = trim(pa, threshold=self.scale * 2.4 + self.shift) if do_trim else 0
        self.template_start = self.trimmed_samples
        self.signal = ((pa[self.trimmed_samples:] - self.shift) / self.scale).astype(np.float32)
        self.scaling_strategy = (scaling_strategy or {}).get("strategy") or "quantile"
        # acquisition metadata the SAM tags carry (bonito/reader.py:59-87); .npy reads have none
        self.run_id, self.mux, self.channel, self.read_number = meta.get("run_id", "unknown"), meta.get("mux", 0), \
            meta.get("channel", 0), meta.get("read_number", 0)
        self.start_time, self.duration = meta.get("start_time", ""), meta.get("duration", self.num_samples / 5000.0)
        self.flow_cell_id, self.device_id, self.sample_id, self.exp_start_time = (
            meta.get(k, "") for k in ("flow_cell_id", "device_id", "sample_id", "exp_start_time"))

    def readgroup(self, model):
        """@RG header line (reference: bonito/reader.py:59-73)."""
        fields = [("ID", f"{self.run_id}_{model}"), ("PL", "ONT"), ("DT", self.exp_start_time), ("PU", self.flow_cell_id),
                  ("PM", self.device_id), ("LB", self.sample_id), ("SM", self.sample_id),
                  ("DS", f"run_id={self.run_id} basecall_model={model}")]
        return "\t".join(["@RG", *[f"{k}:{v}" for k, v in fields]])

    def tagdata(self):
        """Per-read SAM tags (reference: bonito/reader.py:75-86)."""
        return [f"mx:i:{self.mux}", f"ch:i:{self.channel}", f"st:Z:{self.start_time}", f"du:f:{self.duration}",
                f"rn:i:{self.read_number}", f"f5:Z:{self.filename}", f"sm:f:{self.shift}", f"sd:f:{self.scale}",
                f"sv:Z:{self.scaling_strategy}"]


class Reader:
    def __init__(self, directory, recursive=False):
        self.fmt = None
        for fmt in ("pod5", "npy"):
            pattern = f"**/*.{fmt}" if recursive else f"*.{fmt}"
            if glob(os.path.join(directory, pattern), recursive=True):
                self.fmt = fmt
                break
        if self.fmt is None:
            raise FileNotFoundError(directory)
        if self.fmt == "pod5":
            try:
                import pod5  # noqa: F401
            except ImportError as err:
                raise FileNotFoundError(f"{directory}: pod5 files found but the `pod5` package is not installed") from err

    def get_reads(self, directory, recursive=False, read_ids=None, skip=False, do_trim=True, scaling_strategy=None,
                  norm_params=None, **_ignored):
        pattern = f"**/*.{self.fmt}" if recursive else f"*.{self.fmt}"
        for path in sorted(glob(os.path.join(directory, pattern), recursive=True)):
            for read_id, pa in self._signals(path):
                if read_ids is not None and ((read_id in read_ids) == bool(skip)):
                    continue
                yield Read(read_id, pa, filename=os.path.basename(path), do_trim=do_trim,
                           scaling_strategy=scaling_strategy, norm_params=norm_params)

    def _signals(self, path):
        if self.fmt == "npy":
            yield os.path.splitext(os.path.basename(path))[0], np.load(path)
            return
        import pod5
        with pod5.Reader(path) as reader:
            for rec in reader.reads():
                cal = rec.calibration
                yield str(rec.read_id), cal.scale * (rec.signal.astype(np.float32) + cal.offset)
