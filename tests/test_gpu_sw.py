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


def test_batch_api_edge_cases(sw):
    """Empty batch is a no-op, an unknown strategy code is refused (IntelSmithWaterman.java:144), one long and one
    one-base pair share a batch (scratch is sized by the largest pair)."""
    import ctypes as C
    from gkl_b200 import native
    from gkl_b200.smithwaterman import _SwBatch, _lib
    lib = _lib()
    empty = _SwBatch(0, None, None, None, None, 3, -1, -4, -3, 9)
    assert lib.gklb_sw_align_batch(C.byref(empty), C.c_void_p(1), 8, C.c_void_p(1), C.c_void_p(1)) == native.OK
    s1, o1 = pack([b"ACGT"])
    s2, o2 = pack([b"ACGT"])
    cig = np.zeros((1, 8), dtype=np.uint8)
    clen, offs = np.zeros(1, np.int32), np.zeros(1, np.int32)
    bad = _SwBatch(1, s1.ctypes.data, o1.ctypes.data, s2.ctypes.data, o2.ctypes.data, 3, -1, -4, -3, 13)
    assert lib.gklb_sw_align_batch(C.byref(bad), cig.ctypes.data, 8, clen.ctypes.data, offs.ctypes.data) == native.ERR_INVALID
    rng = np.random.default_rng(3)
    long_a = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=5000)]
    long_b = np.delete(long_a, slice(1000, 1040))
    long_b[2500] = ord("A") if long_b[2500] != ord("A") else ord("C")
    a, b = [long_a.tobytes(), b"G"], [long_b.tobytes(), b"G"]
    s1, o1 = pack(a)
    s2, o2 = pack(b)
    for strat in STRATEGIES:
        ref = oracle.port_sw(s1, o1, s2, o2, PARAMS[0], strat, threads=2)
        cigs, off = sw.align_batch(a, b, SWParameters(*PARAMS[0]), SWOverhangStrategy(strat))
        assert cigs == ref[0] and np.array_equal(off, ref[1])
    assert cigs[1] == "1M"


def test_cigar_buffer_rule_when_the_caller_buffer_is_short(sw):
    """getCIGAR drops the elements that do not fit into what is left of the caller's buffer (PairWiseSW.h:426-431); the
    one-pair entry point takes the buffer length like runSWOnePairBT does."""
    import ctypes as C
    from gkl_b200.smithwaterman import _lib
    ref = np.frombuffer(b"ACGTACGTTTGACCA" * 4, dtype=np.uint8).copy()
    alt = np.concatenate([ref[:20], ref[26:]]).copy()
    full = sw.align(ref.tobytes(), alt.tobytes(), SWParameters(3, -1, -4, -3), SWOverhangStrategy.SOFTCLIP).cigar
    assert "D" in full and len(full) > 5
    for cap in (len(full), len(full) - 1, 3, 2):
        buf = np.zeros(cap, dtype=np.uint8)
        count, off = C.c_uint32(0), C.c_int32(0)
        rc = _lib().gklb_sw_align(3, -1, -4, -3, ref.ctypes.data, alt.ctypes.data, len(ref), len(alt), 9, buf.ctypes.data,
                                  cap, C.byref(count), C.byref(off))
        assert rc == 0
        # the same rule restated: elements first to last, each written only if it still fits
        import re
        want = ""
        for el in re.findall(r"\d+[MIDS]", full):
            if len(want) + len(el) <= cap:
                want += el
        assert bytes(buf[:count.value]).decode() == want


def test_large_batches_are_cut_into_chunks(sw):
    """The run arena of one launch is bounded; with the bound forced low (it is read once per process, so a fresh
    process) the same batch goes through several launches and must give the same answers."""
    import subprocess
    import sys
    code = (
        "import os, sys, numpy as np\n"
        "os.environ['GKLB_SW_MAX_RUN_ELEMENTS'] = '900'\n"
        "sys.path.insert(0, '.')\n"
        "import oracle\n"
        "from gkl_b200.smithwaterman import IntelSmithWaterman, SWOverhangStrategy, SWParameters\n"
        "from tests.test_oracle_sw import pack, random_pairs\n"
        "a, b, params, strat = random_pairs(3, 300, max_len=60)\n"
        "s1, o1 = pack(a); s2, o2 = pack(b)\n"
        "ref = oracle.port_sw(s1, o1, s2, o2, params, strat, threads=2)\n"
        "sw = IntelSmithWaterman(); assert sw.load()\n"
        "cig, off = sw.align_batch(a, b, SWParameters(*params), SWOverhangStrategy(strat))\n"
        "st = sw.stats(); sw.close()\n"
        "assert cig == ref[0] and np.array_equal(off, ref[1])\n"
        "assert st.kernel_launches > 5 and st.pairs == 300, (st.kernel_launches, st.pairs)\n"
        "print('chunks ok', st.kernel_launches)\n")
    from pathlib import Path
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(Path(__file__).resolve().parents[1]))
    assert r.returncode == 0 and "chunks ok" in r.stdout, r.stdout + r.stderr
