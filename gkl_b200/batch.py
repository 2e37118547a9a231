"""Flat read x haplotype batch: the layout the C-ABI (include/gklb_pairhmm.h) consumes.

GKL's JNI glue walks ``ReadDataHolder[]`` / ``HaplotypeDataHolder[]`` and materialises one
``testcase`` per (read, haplotype) pair (reference pairhmm/JavaData.h:65-111,
pairhmm/pairhmm_common.h:43-47).  Here the same information is six byte arenas plus two
offset arrays; a pair is addressed by index arithmetic only (pair ``r * n_haps + h``, the
order of JavaData.h:94-105), so nothing of size reads x haps is ever built on the host.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np


def _arena(seqs: Sequence[bytes | bytearray | np.ndarray]) -> tuple[np.ndarray, np.ndarray]:
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    arena = np.empty(int(off[-1]), dtype=np.uint8)
    for i, s in enumerate(seqs):
        arena[off[i]:off[i + 1]] = np.frombuffer(bytes(s), dtype=np.uint8) if not isinstance(s, np.ndarray) else s
    return off, arena


@dataclass
class PairHmmBatch:
    """Reads (five parallel byte arenas) and haplotypes (one arena), with offsets."""

    read_off: np.ndarray   # int64[n_reads + 1]
    read_bases: np.ndarray  # uint8[total_read_len]  ASCII bases
    read_quals: np.ndarray  # uint8[...]  base qualities (already Phred-decoded, as GATK passes them)
    ins_gop: np.ndarray    # uint8[...]  insertion gap-open penalties
    del_gop: np.ndarray    # uint8[...]  deletion gap-open penalties
    gcp: np.ndarray        # uint8[...]  gap-continuation penalties
    hap_off: np.ndarray    # int64[n_haps + 1]
    hap_bases: np.ndarray  # uint8[total_hap_len]

    @property
    def n_reads(self) -> int:
        return len(self.read_off) - 1

    @property
    def n_haps(self) -> int:
        return len(self.hap_off) - 1

    @property
    def read_lens(self) -> np.ndarray:
        return np.diff(self.read_off)

    @property
    def hap_lens(self) -> np.ndarray:
        return np.diff(self.hap_off)

    def cells(self) -> int:
        """Cell updates in the batch: sum over pairs of rslen x haplen (JavaData.h:108)."""
        return int(self.read_off[-1]) * int(self.hap_off[-1])

    def input_bytes(self) -> int:
        return 5 * int(self.read_off[-1]) + int(self.hap_off[-1])

    @staticmethod
    def from_lists(read_bases, read_quals, ins_gop, del_gop, gcp, hap_bases) -> "PairHmmBatch":
        n = len(read_bases)
        if not (len(read_quals) == len(ins_gop) == len(del_gop) == len(gcp) == n):
            raise ValueError("per-read lists must have equal length")
        read_off, rb = _arena(read_bases)
        arenas = []
        for name, lst in (("read_quals", read_quals), ("ins_gop", ins_gop), ("del_gop", del_gop), ("gcp", gcp)):
            off, a = _arena(lst)
            if not np.array_equal(off, read_off):
                raise ValueError(f"{name} lengths differ from read_bases lengths")
            arenas.append(a)
        hap_off, hb = _arena(hap_bases)
        return PairHmmBatch(read_off, rb, arenas[0], arenas[1], arenas[2], arenas[3], hap_off, hb)

    def read_slice(self, lo: int, hi: int) -> "PairHmmBatch":
        """Reads [lo, hi) against all haplotypes (the unit of multi-GPU sharding)."""
        a, b = int(self.read_off[lo]), int(self.read_off[hi])
        return PairHmmBatch(
            (self.read_off[lo:hi + 1] - a).astype(np.int64), self.read_bases[a:b], self.read_quals[a:b],
            self.ins_gop[a:b], self.del_gop[a:b], self.gcp[a:b], self.hap_off, self.hap_bases)

    def validate(self) -> None:
        for name in ("read_bases", "read_quals", "ins_gop", "del_gop", "gcp", "hap_bases"):
            a = getattr(self, name)
            if a.dtype != np.uint8 or not a.flags.c_contiguous:
                raise ValueError(f"{name} must be contiguous uint8")
        for name in ("read_off", "hap_off"):
            a = getattr(self, name)
            if a.dtype != np.int64 or not a.flags.c_contiguous:
                raise ValueError(f"{name} must be contiguous int64")
        n = int(self.read_off[-1])
        for name in ("read_bases", "read_quals", "ins_gop", "del_gop", "gcp"):
            if len(getattr(self, name)) != n:
                raise ValueError(f"{name} has {len(getattr(self, name))} bytes, offsets say {n}")
        if len(self.hap_bases) != int(self.hap_off[-1]):
            raise ValueError("hap_bases length does not match hap_off")
