"""Generates tests/golden/sw_golden.txt.gz (run in the development container, where /root/reference exists).

Inputs: every 47th pair of the reference's src/test/resources/smith-waterman.SOFTCLIP.in (its own test data; the
reference's SmithWatermanUnitTest.simpleTest aligns them but asserts nothing), plus the two known-answer pairs of
SmithWatermanUnitTest.java:160-190.  Expected outputs: GKL's own compiled Smith-Waterman (oracle/_ref, gklref_sw) for
every overhang strategy under the reference test's parameters (200,-150,-260,-11) and its commented-out set
(3,-1,-4,-3) (SmithWatermanUnitTest.java:45-46).

Line format: ref <TAB> alt <TAB> cigar:offset x 8  (parameter set major, strategies 9..12 minor).
"""
import gzip
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402

PARAMS = [(200, -150, -260, -11), (3, -1, -4, -3)]
STRATEGIES = [9, 10, 11, 12]


def pack(seqs):
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    return np.frombuffer("".join(seqs).encode(), dtype=np.uint8).copy(), off


def main():
    lines = [l for l in open("/root/reference/src/test/resources/smith-waterman.SOFTCLIP.in").read().split("\n") if l]
    refs, alts = lines[0::2], lines[1::2]
    n = min(len(refs), len(alts))
    pick = list(range(0, n, 47))
    refs = [refs[k] for k in pick] + ["C", "AD"]
    alts = [alts[k] for k in pick] + ["C", "AT"]
    s1, o1 = pack(refs)
    s2, o2 = pack(alts)
    cols = []
    for p in PARAMS:
        for st in STRATEGIES:
            cig, off, _ = oracle.ref_sw(s1, o1, s2, o2, p, st, threads=oracle.host_threads())
            cols.append([f"{c}:{o}" for c, o in zip(cig, off)])
    with gzip.open(Path(__file__).with_name("sw_golden.txt.gz"), "wt") as f:
        for k in range(len(refs)):
            f.write("\t".join([refs[k], alts[k]] + [c[k] for c in cols]) + "\n")
    print(len(refs), "pairs")


if __name__ == "__main__":
    main()
