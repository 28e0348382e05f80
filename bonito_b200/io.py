"""
pysam-free FASTQ output for `bonito_b200 basecaller` (the reference's `Writer` thread, `/root/reference/bonito/io.py:400-469`,
needs pysam / mappy; unaligned FASTQ is the format it emits on a terminal or a `.fastq` redirect, `io.py:35-54`).
"""

import sys
from threading import Thread

from bonito_b200.util import mean_qscore_from_qstring


def write_fastq(header, sequence, qstring, fd=sys.stdout):
    fd.write(f"@{header}\n{sequence}\n+\n{qstring}\n")


class Writer(Thread):
    """Drains the basecall iterator on its own thread; `.log` holds (read_id, num_samples) of the reads written."""

    def __init__(self, iterator, fd=sys.stdout, min_qscore=0):
        super().__init__(daemon=True)
        self.iterator, self.fd, self.min_qscore = iterator, fd, min_qscore
        self.log, self.error = [], None

    def run(self):
        try:
            for read, res in self.iterator:
                seq, qstring = res["sequence"], res["qstring"]
                samples = len(read.signal) + getattr(read, "trimmed_samples", 0)
                if len(seq) and mean_qscore_from_qstring(qstring) >= self.min_qscore:
                    write_fastq(read.read_id, seq, qstring, fd=self.fd)
                    self.log.append((read.read_id, samples))
                else:
                    sys.stderr.write(f"> skipping empty / low quality sequence {read.read_id}\n")
            self.fd.flush()
        except BaseException as err:   # surfaced by the CLI after join()
            self.error = err
