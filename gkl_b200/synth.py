"""Deterministic synthetic read x haplotype batches for the BASELINE.json configurations.

Shapes and distributions follow SURVEY.md section 8(d).  The quality histograms are the empirical
ones derived there from the real reads of GKL's src/test/resources/pdhmm_new.txt.  The generator
is numpy's PCG64 seeded per configuration, so every run (here, on the GPU box, in tests and in
bench.py) sees byte-identical inputs.
"""
from __future__ import annotations

import numpy as np

from .batch import PairHmmBatch

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)

# (value, probability) -- SURVEY.md 8(d) "Empirical qual distribution"
H_Q = {6: .015, 18: .003, 19: .003, 20: .003, 21: .005, 22: .006, 23: .006, 24: .007, 25: .009, 26: .010,
       27: .013, 28: .015, 29: .019, 30: .018, 31: .020, 32: .027, 33: .034, 34: .037, 35: .061, 36: .086,
       37: .190, 38: .270, 39: .136, 40: .007, 41: .001}
H_GOP = {22: .069, 23: .006, 24: .003, 25: .001, 26: .001, 28: .076, 29: .007, 32: .005, 33: .002, 37: .023,
         38: .007, 40: .786, 45: .014}
H_GCP = {1: .007, 2: .023, 3: .120, 5: .117, 10: .733}


def _draw(rng: np.random.Generator, hist: dict[int, float], n: int) -> np.ndarray:
    vals = np.fromiter(hist.keys(), dtype=np.uint8)
    p = np.fromiter(hist.values(), dtype=np.float64)
    return rng.choice(vals, size=n, p=p / p.sum())


def _haplotype_panel(rng, n_haps: int, lo: int, hi: int) -> list[np.ndarray]:
    """One random ancestor; each haplotype is a prefix of it with 2 % SNPs and, with probability
    0.5, one 1-10 base indel -- a panel that looks like an assembled active region."""
    ancestor = ACGT[rng.integers(0, 4, size=hi)]
    haps = []
    for _ in range(n_haps):
        L = int(rng.integers(lo, hi + 1))
        h = ancestor[:L].copy()
        snp = rng.random(L) < 0.02
        h[snp] = ACGT[rng.integers(0, 4, size=int(snp.sum()))]
        if rng.random() < 0.5 and L > 30:
            pos = int(rng.integers(10, L - 10))
            k = int(rng.integers(1, 11))
            if rng.random() < 0.5:  # deletion, then re-extend from the ancestor so the length stays L
                h = np.concatenate([h[:pos], h[pos + k:], ACGT[rng.integers(0, 4, size=k)]])
            else:  # insertion, trimmed back to L
                h = np.concatenate([h[:pos], ACGT[rng.integers(0, 4, size=k)], h[pos:]])[:L]
        haps.append(np.ascontiguousarray(h))
    return haps


def _reads_from_panel(rng, haps: list[np.ndarray], read_lens: np.ndarray, empirical: bool):
    n = len(read_lens)
    total = int(read_lens.sum())
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(read_lens, out=off[1:])
    bases = np.empty(total, dtype=np.uint8)
    src = rng.integers(0, len(haps), size=n)
    start_u = rng.random(n)
    for r in range(n):
        h = haps[src[r]]
        L = int(read_lens[r])
        if len(h) >= L:
            s = int(start_u[r] * (len(h) - L + 1))
            seg = h[s:s + L]
        else:  # read longer than its source haplotype: pad with random bases
            seg = np.concatenate([h, ACGT[rng.integers(0, 4, size=L - len(h))]])
        bases[off[r]:off[r + 1]] = seg
    sub = rng.random(total) < 0.01
    bases[sub] = ACGT[rng.integers(0, 4, size=int(sub.sum()))]
    bases[rng.random(total) < 0.005] = ord("N")
    if empirical:
        quals = _draw(rng, H_Q, total)
        gop = _draw(rng, H_GOP, total)  # insertion == deletion GOP, drawn once per base
        gcp = _draw(rng, H_GCP, total)
    else:
        quals = np.full(total, 30, dtype=np.uint8)
        gop = np.full(total, 40, dtype=np.uint8)
        gcp = np.full(total, 10, dtype=np.uint8)
    return off, bases, quals, gop, gop.copy(), gcp


def _assemble(haps, reads) -> PairHmmBatch:
    hap_lens = np.fromiter((len(h) for h in haps), dtype=np.int64, count=len(haps))
    hap_off = np.zeros(len(haps) + 1, dtype=np.int64)
    np.cumsum(hap_lens, out=hap_off[1:])
    off, bases, quals, ins, dele, gcp = reads
    return PairHmmBatch(off, bases, quals, ins, dele, gcp, hap_off, np.concatenate(haps))


def config1() -> PairHmmBatch:
    """C1: 1 read (len 50) x 1 haplotype (len 100), uniform Q30, GOP 40, GCP 10."""
    rng = np.random.default_rng(1)
    hap = ACGT[rng.integers(0, 4, size=100)]
    read = hap[25:75].copy()
    n = 50
    return PairHmmBatch(np.array([0, n], dtype=np.int64), read, np.full(n, 30, np.uint8), np.full(n, 40, np.uint8),
                        np.full(n, 40, np.uint8), np.full(n, 10, np.uint8), np.array([0, 100], dtype=np.int64), hap)


def config2(n_reads: int = 10_000, n_haps: int = 128, read_len: int = 101, seed: int = 2) -> PairHmmBatch:
    """C2 (the headline configuration): 10 000 reads (len 101) x 128 haplotypes (len 200-400),
    empirical quality distribution."""
    rng = np.random.default_rng(seed)
    haps = _haplotype_panel(rng, n_haps, 200, 400)
    reads = _reads_from_panel(rng, haps, np.full(n_reads, read_len, dtype=np.int64), empirical=True)
    return _assemble(haps, reads)


def config3(n_regions: int = 32, seed: int = 3) -> list[PairHmmBatch]:
    """C3: HaplotypeCaller-shaped -- 32 independent active regions, ~300 reads x ~64 haplotypes each,
    read lengths 35-250; one synchronous call per region."""
    rng = np.random.default_rng(seed)
    regions = []
    for _ in range(n_regions):
        R = int(rng.integers(200, 401))
        H = int(rng.integers(32, 97))
        lens = rng.integers(35, 251, size=R).astype(np.int64)
        hap_len = int(lens.max()) + int(rng.integers(50, 201))
        haps = _haplotype_panel(rng, H, hap_len - 20, hap_len)
        regions.append(_assemble(haps, _reads_from_panel(rng, haps, lens, empirical=True)))
    return regions


def config4(n_reads: int = 1_000_000, n_haps: int = 256, read_len: int = 150, seed: int = 4) -> PairHmmBatch:
    """C4: 1 M reads (len 150) x 256 haplotypes; sharded over reads across GPUs."""
    rng = np.random.default_rng(seed)
    haps = _haplotype_panel(rng, n_haps, 200, 400)
    reads = _reads_from_panel(rng, haps, np.full(n_reads, read_len, dtype=np.int64), empirical=True)
    return _assemble(haps, reads)


def config5(n_reads: int = 10_000, n_haps: int = 128, read_len: int = 101, seed: int = 5):
    """C5: the PDHMM variant of C2 -- same shapes, plus one PD flag byte per haplotype base drawn from the
    frequencies of the reference's real-data fixture: 97.2 % plain, 1.8 % SNP|G (33), 0.3 % SNP|A (9),
    and DEL_START (2) ... DEL_END (4) spans opened at 0.4 % of the positions (1-8 bases; a 1-base span carries both
    flags).  Returns (reads, haps) for gkl_b200.pdhmm_batch.PdhmmBatch.cross / IntelPDHMM.computeLikelihoods:
    reads = [(bases, qual, ins, del, gcp)], haps = [(bases, pd)], all int8."""
    b = config2(n_reads, n_haps, read_len, seed=seed)
    rng = np.random.default_rng(seed + 1000)
    haps = []
    for h in range(b.n_haps):
        bases = b.hap_bases[b.hap_off[h]:b.hap_off[h + 1]]
        L = len(bases)
        pd = np.zeros(L, dtype=np.int8)
        u = rng.random(L)
        pd[u < 0.018] = 33
        pd[(u >= 0.018) & (u < 0.021)] = 9
        for pos in np.flatnonzero(rng.random(L) < 0.004):
            span = int(rng.integers(1, 9))
            end = min(L - 1, pos + span - 1)
            pd[pos] |= 2
            pd[end] |= 4
        haps.append((bases.view(np.int8).copy(), pd))
    reads = []
    for r in range(b.n_reads):
        a, e = int(b.read_off[r]), int(b.read_off[r + 1])
        reads.append(tuple(x[a:e].view(np.int8).copy() for x in (b.read_bases, b.read_quals, b.ins_gop, b.del_gop, b.gcp)))
    return reads, haps


def random_batch(seed: int, n_reads: int, n_haps: int, read_len=(10, 250), hap_len=(10, 450),
                 low_quality: float = 0.0, n_frac: float = 0.02, unrelated: float = 0.0) -> PairHmmBatch:
    """Adversarial parity batch: ragged lengths, N and non-ACGT bytes on both sides, a share of
    low GOP/GCP values, and (``unrelated``) reads that do not come from any haplotype, which
    drives the fp32 sum under GKL's 1e-28 threshold and into the fp64 rerun."""
    rng = np.random.default_rng(seed)
    haps = [ACGT[rng.integers(0, 4, size=int(rng.integers(hap_len[0], hap_len[1] + 1)))] for _ in range(n_haps)]
    for h in haps:
        h[rng.random(len(h)) < n_frac] = ord("N")
        h[rng.random(len(h)) < n_frac / 4] = ord("a")  # anything but ACGTN is treated as 'A'
    lens = rng.integers(read_len[0], read_len[1] + 1, size=n_reads).astype(np.int64)
    off, bases, quals, ins, dele, gcp = _reads_from_panel(rng, haps, lens, empirical=True)
    total = len(bases)
    if unrelated > 0:
        for r in range(n_reads):
            if rng.random() < unrelated:
                bases[off[r]:off[r + 1]] = ACGT[rng.integers(0, 4, size=int(lens[r]))]
                quals[off[r]:off[r + 1]] = rng.integers(35, 94, size=int(lens[r]))
    if low_quality > 0:
        lo = rng.random(total) < low_quality
        quals[lo] = rng.integers(0, 128, size=int(lo.sum()))
        lo = rng.random(total) < low_quality
        ins[lo] = rng.integers(0, 128, size=int(lo.sum()))
        lo = rng.random(total) < low_quality
        dele[lo] = rng.integers(0, 128, size=int(lo.sum()))
        lo = rng.random(total) < low_quality
        gcp[lo] = rng.integers(0, 128, size=int(lo.sum()))
    bases[rng.random(total) < n_frac / 4] = ord("c")
    return _assemble(haps, (off, bases, quals, ins, dele, gcp))
