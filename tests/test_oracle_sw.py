"""CPU: the Smith-Waterman restatement (oracle/sw_oracle.c) against the golden file produced by GKL's own compiled
code, the reference's two known answers, and -- when /root/reference is present -- GKL's code itself on random
adversarial pairs (small alphabets: many equal maxima)."""
import gzip

import numpy as np
import pytest

import oracle
from tests.conftest import GOLDEN

PARAMS = [(200, -150, -260, -11), (3, -1, -4, -3)]
STRATEGIES = [9, 10, 11, 12]


def load_sw_golden():
    refs, alts, exp = [], [], []
    with gzip.open(GOLDEN / "sw_golden.txt.gz", "rt") as f:
        for line in f:
            sp = line.rstrip("\n").split("\t")
            refs.append(sp[0].encode())
            alts.append(sp[1].encode())
            exp.append([(x.rsplit(":", 1)[0], int(x.rsplit(":", 1)[1])) for x in sp[2:]])
    return refs, alts, exp


def pack(seqs):
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    return np.frombuffer(b"".join(seqs), dtype=np.uint8).copy(), off


def random_pairs(seed, n, max_len=30):
    rng = np.random.default_rng(seed)
    alpha = int(rng.integers(1, 5))
    a, b = [], []
    for _ in range(n):
        x = rng.integers(65, 65 + alpha, size=int(rng.integers(1, max_len))).astype(np.uint8)
        if rng.random() < 0.5:
            y = x.copy()
            for _ in range(int(rng.integers(0, 4))):
                if len(y) > 1 and rng.random() < 0.5:
                    y = np.delete(y, int(rng.integers(0, len(y))))
                else:
                    y = np.insert(y, int(rng.integers(0, len(y) + 1)), int(rng.integers(65, 65 + alpha)))
            y = y.astype(np.uint8)
        else:
            y = rng.integers(65, 65 + alpha, size=int(rng.integers(1, max_len))).astype(np.uint8)
        a.append(x.tobytes())
        b.append(y.tobytes())
    params = (int(rng.integers(1, 30)), -int(rng.integers(0, 30)), -int(rng.integers(0, 40)), -int(rng.integers(0, 12)))
    return a, b, params, int(rng.integers(9, 13))


def test_restatement_matches_the_golden_file():
    refs, alts, exp = load_sw_golden()
    s1, o1 = pack(refs)
    s2, o2 = pack(alts)
    col = 0
    for p in PARAMS:
        for st in STRATEGIES:
            cig, off, _ = oracle.port_sw(s1, o1, s2, o2, p, st, threads=oracle.host_threads())
            for k in range(len(refs)):
                assert (cig[k], int(off[k])) == exp[k][col], (p, st, k)
            col += 1


def test_known_answers_of_the_reference_tests():
    # SmithWatermanUnitTest.java:160-190: IGNORE, (3,-2,-2,-1) "C" vs "C" -> 1M; (3,-5,-2,-1) "AD" vs "AT" -> 1M1I
    s1, o1 = pack([b"C"])
    s2, o2 = pack([b"C"])
    assert oracle.port_sw(s1, o1, s2, o2, (3, -2, -2, -1), 12)[0] == ["1M"]
    s1, o1 = pack([b"AD"])
    s2, o2 = pack([b"AT"])
    assert oracle.port_sw(s1, o1, s2, o2, (3, -5, -2, -1), 12)[0] == ["1M1I"]


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_restatement_matches_gkl_on_adversarial_pairs():
    for seed in range(12):
        a, b, params, strat = random_pairs(seed, 1500)
        s1, o1 = pack(a)
        s2, o2 = pack(b)
        p = oracle.port_sw(s1, o1, s2, o2, params, strat, threads=oracle.host_threads())
        for engine in (1, 2) if oracle.ref_avx512_supported() else (1,):
            r = oracle.ref_sw(s1, o1, s2, o2, params, strat, threads=oracle.host_threads(), engine=engine)
            assert p[0] == r[0] and np.array_equal(p[1], r[1]), (seed, engine)


@pytest.mark.skipif(not oracle.ref_available() or not oracle.REFERENCE_ROOT.is_dir(),
                    reason="needs /root/reference and oracle/_ref (development container)")
def test_restatement_matches_gkl_on_the_whole_reference_test_file():
    """All 14 196 pairs of src/test/resources/smith-waterman.SOFTCLIP.in with the parameters and strategy of
    SmithWatermanUnitTest.simpleTest (which aligns them without asserting anything)."""
    path = oracle.REFERENCE_ROOT / "src/test/resources/smith-waterman.SOFTCLIP.in"
    lines = [l for l in path.read_text().split("\n") if l]
    refs = [l.encode() for l in lines[0::2]]
    alts = [l.encode() for l in lines[1::2]]
    n = min(len(refs), len(alts))
    s1, o1 = pack(refs[:n])
    s2, o2 = pack(alts[:n])
    threads = oracle.host_threads()
    p = oracle.port_sw(s1, o1, s2, o2, PARAMS[0], 9, threads=threads)
    r = oracle.ref_sw(s1, o1, s2, o2, PARAMS[0], 9, threads=threads)
    assert n == 14196 and p[0] == r[0] and np.array_equal(p[1], r[1])
