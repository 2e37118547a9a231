"""Flat PDHMM batch: the layout of IntelPDHMM.computePDHMM (reference
src/main/java/com/intel/gkl/pdhmm/IntelPDHMM.java:163-204 and pdhmm/pdhmm-common.h:78-136).

Pair k uses hap_bases / hap_pdbases[k * max_hap : k * max_hap + hap_lengths[k]] and the five read arrays
[k * max_read : k * max_read + read_lengths[k]]; everything past the lengths is zero padding.  The object
API (IntelPDHMM.computeLikelihoods, :92-121) expands reads x haplotypes read-major into this layout
(pdhmm/JavaData.h:177-242); `cross()` does the same expansion.
"""
from __future__ import annotations

import gzip
from dataclasses import dataclass
from pathlib import Path

import numpy as np


@dataclass
class PdhmmBatch:
    hap_bases: np.ndarray      # int8[n * max_hap]
    hap_pdbases: np.ndarray    # int8[n * max_hap]  PD flag bytes: SNP=1 DEL_START=2 DEL_END=4 A=8 C=16 G=32 T=64
    read_bases: np.ndarray     # int8[n * max_read]
    read_qual: np.ndarray
    read_ins_qual: np.ndarray
    read_del_qual: np.ndarray
    gcp: np.ndarray
    hap_lengths: np.ndarray    # int64[n]
    read_lengths: np.ndarray   # int64[n]
    max_hap: int
    max_read: int

    @property
    def n(self) -> int:
        return len(self.hap_lengths)

    def cells(self) -> int:
        return int(np.sum(self.hap_lengths * self.read_lengths))

    def slice(self, lo: int, hi: int) -> "PdhmmBatch":
        h, r = self.max_hap, self.max_read
        return PdhmmBatch(self.hap_bases[lo * h:hi * h], self.hap_pdbases[lo * h:hi * h], self.read_bases[lo * r:hi * r],
                          self.read_qual[lo * r:hi * r], self.read_ins_qual[lo * r:hi * r],
                          self.read_del_qual[lo * r:hi * r], self.gcp[lo * r:hi * r], self.hap_lengths[lo:hi],
                          self.read_lengths[lo:hi], h, r)

    @staticmethod
    def from_pairs(pairs) -> "PdhmmBatch":
        """pairs: iterable of (hap, pd, read, qual, ins, del, gcp) byte-like / int8 arrays."""
        pairs = [tuple(np.frombuffer(bytes(x), dtype=np.int8) if not isinstance(x, np.ndarray) else x.astype(np.int8)
                       for x in p) for p in pairs]
        n = len(pairs)
        max_hap = max((len(p[0]) for p in pairs), default=1)
        max_read = max((len(p[2]) for p in pairs), default=1)
        hb, pd = (np.zeros(n * max_hap, dtype=np.int8) for _ in range(2))
        rb, q, i, d, c = (np.zeros(n * max_read, dtype=np.int8) for _ in range(5))
        hl, rl = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
        for k, (h, p, r, qq, ii, dd, cc) in enumerate(pairs):
            hl[k], rl[k] = len(h), len(r)
            hb[k * max_hap:k * max_hap + len(h)] = h
            pd[k * max_hap:k * max_hap + len(p)] = p
            for dst, src in ((rb, r), (q, qq), (i, ii), (d, dd), (c, cc)):
                dst[k * max_read:k * max_read + len(src)] = src
        return PdhmmBatch(hb, pd, rb, q, i, d, c, hl, rl, max_hap, max_read)

    @staticmethod
    def operands(reads, haps) -> "PdhmmBatch":
        """The operands of IntelPDHMM.computeLikelihoods once each (what gklb_pdhmm_compute_cross takes): haplotype h
        at h * max_hap (hap_lengths[H]), read r at r * max_read (read_lengths[R]); the cross product is device-side."""
        ops = PdhmmBatch.from_pairs([(h[0], h[1], b"", b"", b"", b"", b"") for h in haps])
        rds = PdhmmBatch.from_pairs([(b"", b"", r[0], r[1], r[2], r[3], r[4]) for r in reads])
        return PdhmmBatch(ops.hap_bases, ops.hap_pdbases, rds.read_bases, rds.read_qual, rds.read_ins_qual,
                          rds.read_del_qual, rds.gcp, ops.hap_lengths, rds.read_lengths, ops.max_hap, rds.max_read)

    def expand_cross(self, r0: int, r1: int) -> "PdhmmBatch":
        """Flat batch of reads [r0, r1) x all haplotypes of an `operands` object, pair index (r - r0) * H + h
        (the expansion of pdhmm/JavaData.h:177-242, vectorised)."""
        H, nr = len(self.hap_lengths), r1 - r0
        hap = lambda a: np.tile(a.reshape(H, self.max_hap), (nr, 1)).ravel()
        rd = lambda a: np.repeat(a.reshape(-1, self.max_read)[r0:r1], H, axis=0).ravel()
        return PdhmmBatch(hap(self.hap_bases), hap(self.hap_pdbases), rd(self.read_bases), rd(self.read_qual),
                          rd(self.read_ins_qual), rd(self.read_del_qual), rd(self.gcp), np.tile(self.hap_lengths, nr),
                          np.repeat(self.read_lengths[r0:r1], H), self.max_hap, self.max_read)

    @staticmethod
    def cross(reads, haps) -> "PdhmmBatch":
        """reads: list of (bases, qual, ins, del, gcp); haps: list of (bases, pd).  Pair index r * H + h."""
        return PdhmmBatch.from_pairs([(h[0], h[1], r[0], r[1], r[2], r[3], r[4]) for r in reads for h in haps])


def _phred(s: str) -> np.ndarray:  # SAMUtils.fastqToPhred
    return (np.frombuffer(s.encode("latin-1"), dtype=np.uint8).astype(np.int16) - 33).astype(np.int8)


def _pd(s: str) -> np.ndarray:
    body = s.strip()[1:-1].strip()
    return np.array([int(v) for v in body.split(",")], dtype=np.int8) if body else np.zeros(0, dtype=np.int8)


def _open(path):
    path = Path(path)
    return gzip.open(path, "rt", encoding="latin-1") if path.suffix == ".gz" else open(path, "rt", encoding="latin-1")


def load_pdhmm_pairs_file(path, limit: int | None = None):
    """Golden files of IntelPDHMMUnitTest.pdhmmPerformanceTest (IntelPDHMMUnitTest.java:161-257):
    hap, [pd bytes], read, qual, insQ, delQ, gcp, expected -- tab separated, Phred+33.  Returns (batch, expected)."""
    pairs, expected = [], []
    with _open(path) as f:
        f.readline()
        for line in f:
            sp = line.rstrip("\n").split("\t")
            if len(sp) < 8:
                continue
            pairs.append((np.frombuffer(sp[0].encode("latin-1"), dtype=np.int8), _pd(sp[1]),
                          np.frombuffer(sp[2].encode("latin-1"), dtype=np.int8), _phred(sp[3]), _phred(sp[4]),
                          _phred(sp[5]), _phred(sp[6])))
            expected.append(float(sp[7]))
            if limit and len(pairs) >= limit:
                break
    return PdhmmBatch.from_pairs(pairs), np.asarray(expected)


def load_pdhmm_new(path):
    """pdhmm_new.txt (IntelPDHMMUnitTest.newPDHMMTest, :446-555): a reads section, a haplotypes section and
    R * H expected values in r * H + h order.  Returns (reads, haps, expected)."""
    reads, haps, expected, section = [], [], [], 0
    with _open(path) as f:
        for line in f:
            if line.startswith("#"):
                section += 1
                continue
            sp = line.rstrip("\n").split("\t")
            if section == 1 and len(sp) >= 5:
                reads.append((np.frombuffer(sp[0].encode("latin-1"), dtype=np.int8), _phred(sp[1]), _phred(sp[2]),
                              _phred(sp[3]), _phred(sp[4])))
            elif section == 2 and len(sp) >= 2:
                haps.append((np.frombuffer(sp[0].encode("latin-1"), dtype=np.int8), _pd(sp[1])))
            elif section == 3 and sp[0].strip():
                expected.append(float(sp[0]))
    return reads, haps, np.asarray(expected)
