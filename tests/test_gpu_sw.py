"""GPU: the Smith-Waterman CUDA path through the C-ABI -- bit-exact CIGARs and offsets against the golden file made
by GKL's own code, against the CPU restatement on adversarial and long (multi-pass) pairs, and the argument checks of
the Java wrapper."""
import numpy as np
import pytest

import oracle
from gkl_b200.pairhmm import IllegalArgumentException, NullPointerException
from gkl_b200.smithwaterman import IntelSmithWaterman, SWOverhangStrategy, SWParameters
from tests.test_oracle_sw import PARAMS, STRATEGIES, load_sw_golden, pack, random_pairs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sw():
    s = IntelSmithWaterman()
    assert s.load()
    yield s
    s.close()


def test_golden_file_all_strategies(sw):
    refs, alts, exp = load_sw_golden()
    col = 0
    for p in PARAMS:
        for st in STRATEGIES:
            cig, off = sw.align_batch(refs, alts, SWParameters(*p), SWOverhangStrategy(st))
            for k in range(len(refs)):
                assert (cig[k], int(off[k])) == exp[k][col], (p, st, k)
            col += 1


def test_known_answers_single_pair_api(sw):
    # SmithWatermanUnitTest.java:160-190
    assert sw.align(b"C", b"C", SWParameters(3, -2, -2, -1), SWOverhangStrategy.IGNORE).cigar == "1M"
    assert sw.align(b"AD", b"AT", SWParameters(3, -5, -2, -1), SWOverhangStrategy.IGNORE).cigar == "1M1I"


@pytest.mark.parametrize("seed", range(8))
def test_adversarial_pairs_match_the_restatement(sw, seed):
    a, b, params, strat = random_pairs(seed, 2000)
    s1, o1 = pack(a)
    s2, o2 = pack(b)
    ref = oracle.port_sw(s1, o1, s2, o2, params, strat, threads=oracle.host_threads())
    cig, off = sw.align_batch(a, b, SWParameters(*params), SWOverhangStrategy(strat))
    assert cig == ref[0]
    assert np.array_equal(off, ref[1])


@pytest.mark.parametrize("strat", STRATEGIES)
def test_long_sequences_take_several_passes(sw, strat):
    """Reference sequences beyond 256 rows are swept in passes of 256 rows with a carried bottom row; mixed lengths
    in one batch, including a pair near the reference's first buffer growth (1024)."""
    rng = np.random.default_rng(100 + strat)
    a, b = [], []
    for n1, n2 in [(257, 40), (600, 580), (1030, 1000), (300, 900), (1, 700), (700, 1), (512, 512), (2100, 1500)]:
        x = rng.integers(0, 4, size=n1)
        y = np.concatenate([x[: n1 // 3], rng.integers(0, 4, size=5), x[n1 // 3: n1]])[:n2] if n1 > 10 else rng.integers(0, 4, size=n2)
        if len(y) < n2:
            y = np.concatenate([y, rng.integers(0, 4, size=n2 - len(y))])
        a.append(np.frombuffer(b"ACGT", dtype=np.uint8)[x].tobytes())
        b.append(np.frombuffer(b"ACGT", dtype=np.uint8)[y].tobytes())
    s1, o1 = pack(a)
    s2, o2 = pack(b)
    for params in PARAMS:
        ref = oracle.port_sw(s1, o1, s2, o2, params, strat, threads=oracle.host_threads())
        cig, off = sw.align_batch(a, b, SWParameters(*params), SWOverhangStrategy(strat))
        assert cig == ref[0]
        assert np.array_equal(off, ref[1])


def test_argument_validation_matches_the_java_wrapper(sw):
    p = SWParameters(10, -5, -10, -10)
    with pytest.raises(NullPointerException):
        sw.align(None, b"ACGT", p, SWOverhangStrategy.SOFTCLIP)
    with pytest.raises(NullPointerException):
        sw.align(b"ACGT", None, p, SWOverhangStrategy.SOFTCLIP)
    with pytest.raises(NullPointerException):
        sw.align(b"ACGT", b"ACGT", None, SWOverhangStrategy.SOFTCLIP)
    with pytest.raises(NullPointerException):
        sw.align(b"ACGT", b"ACGT", p, None)
    with pytest.raises(IllegalArgumentException):
        sw.align(b"", b"AC", p, SWOverhangStrategy.IGNORE)
    with pytest.raises(IllegalArgumentException):
        sw.align(b"AC", b"", p, SWOverhangStrategy.IGNORE)
    with pytest.raises(IllegalArgumentException):
        sw.align(b"A" * 32768, b"TCCG", p, SWOverhangStrategy.IGNORE)
    with pytest.raises(IllegalArgumentException):
        sw.align(b"ACCG", b"TCCG", SWParameters(64 * 1024 + 1, -5, -10, -10), SWOverhangStrategy.IGNORE)
