"""Parsers for GKL's language-neutral test resources (committed under tests/golden/).

pairhmm-testdata.txt: ``hap read qual insGOP delGOP gcp expected-log10`` per row, Phred+33,
decoded exactly as PairHmmUnitTest.dataFileTest does (reference
src/test/java/com/intel/gkl/pairhmm/PairHmmUnitTest.java:206-212,313-319): subtract 33, and
clamp read qualities (only) to >= 6.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from .batch import PairHmmBatch


def _normalize(s: str, minimum: int = 0) -> bytes:
    a = np.frombuffer(s.encode("ascii"), dtype=np.uint8).astype(np.int16) - 33
    a = np.maximum(a, minimum)
    return a.astype(np.int8).tobytes()


def load_pairhmm_testdata(path: str | Path):
    """Returns (list of single-pair PairHmmBatch, expected[np.float64])."""
    batches, expected = [], []
    for line in Path(path).read_text().splitlines():
        if not line.strip() or line.startswith("#"):
            continue
        hap, read, q, i, d, c, exp = line.split()
        batches.append(PairHmmBatch.from_lists(
            [read.encode()], [_normalize(q, 6)], [_normalize(i)], [_normalize(d)], [_normalize(c)], [hap.encode()]))
        expected.append(float(exp))
    return batches, np.asarray(expected, dtype=np.float64)


def simple_test_batch() -> tuple[PairHmmBatch, float]:
    """PairHmmUnitTest.simpleTest (PairHmmUnitTest.java:55-89): un-normalised '+' (=43) quals."""
    plus = b"++++"
    return PairHmmBatch.from_lists([b"ACGT"], [plus], [plus], [plus], [plus], [b"ACGT"]), -6.022797e-01
