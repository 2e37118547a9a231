"""CPU: pin the oracle.  The restatement (oracle/pairhmm_oracle.c) must reproduce every golden vector of
GKL's own PairHMM tests, and agree with GKL's own compiled AVX code (oracle/_ref) where that is built."""
import numpy as np
import pytest

import oracle
from gkl_b200 import fixtures, synth

TOL_GOLDEN = 1e-5  # PairHmmUnitTest.java:88,221 (absolute, on log10 likelihood)


def _run(fn, batches, use_double):
    return np.array([fn(b, use_double)[0][0] for b in batches])


@pytest.mark.parametrize("use_double", [False, True])
def test_port_matches_gkl_golden_file(golden_pairhmm, use_double):
    batches, expected = golden_pairhmm
    assert len(batches) == 104
    got = _run(oracle.port_pairhmm, batches, use_double)
    assert np.abs(got - expected).max() <= TOL_GOLDEN


def test_port_simple_test_known_answer():
    b, expected = fixtures.simple_test_batch()
    for use_double in (False, True):
        assert abs(oracle.port_pairhmm(b, use_double)[0][0] - expected) <= TOL_GOLDEN


needs_ref = pytest.mark.skipif(not oracle.ref_available() and not oracle.REFERENCE_ROOT.is_dir(),
                               reason="oracle/_ref not built and /root/reference absent")


@needs_ref
@pytest.mark.parametrize("use_double", [False, True])
def test_reference_build_matches_golden_file(golden_pairhmm, use_double):
    batches, expected = golden_pairhmm
    got = _run(oracle.ref_pairhmm, batches, use_double)
    assert np.abs(got - expected).max() <= TOL_GOLDEN


@needs_ref
def test_port_agrees_with_reference_build_on_random_batches():
    for seed, kw in ((21, {}), (22, dict(low_quality=0.1, unrelated=0.3)), (23, dict(read_len=(1, 40), hap_len=(1, 60)))):
        b = synth.random_batch(seed, 60, 25, **kw)
        ref = oracle.ref_pairhmm(b, threads=4)[0]
        port, fell_back, _ = oracle.port_pairhmm(b, threads=4)
        assert np.all(np.isfinite(port))
        rel = np.abs(port - ref) / np.abs(ref)
        assert rel.max() <= 1e-5, (seed, rel.max())
        if kw.get("unrelated"):
            assert fell_back.sum() > 0  # the fp64 rerun of IntelPairHmm.cc:159-162 is exercised


@needs_ref
def test_avx_and_avx512_reference_engines_agree():
    b = synth.config2(40, 16)
    a = oracle.ref_pairhmm(b, engine=1)[0]
    if oracle.ref_avx512_supported():
        c = oracle.ref_pairhmm(b, engine=2)[0]
        assert (np.abs(a - c) / np.abs(a)).max() <= 2e-6  # the two engines may contract mul+add differently


def test_output_order_is_read_major():
    # JavaData.h:94-105: index = r * numHaplotypes + h
    b = synth.random_batch(5, 6, 4)
    full = oracle.port_pairhmm(b)[0].reshape(b.n_reads, b.n_haps)
    for r in (0, 3, 5):
        one = oracle.port_pairhmm(b.read_slice(r, r + 1))[0]
        assert np.array_equal(one, full[r])


def test_base_codes_unknown_bytes_are_A_and_N_matches_everything():
    # pairhmm_common.h:53-66 and precompute_masks (avx-pairhmm-template.h:26-58)
    mk = lambda read, hap: fixtures.PairHmmBatch.from_lists([read], [bytes([30]) * len(read)], [bytes([40]) * len(read)],
                                                            [bytes([40]) * len(read)], [bytes([10]) * len(read)], [hap])
    base = oracle.port_pairhmm(mk(b"ACGTAC", b"TTACGTACTT"))[0][0]
    assert oracle.port_pairhmm(mk(b"aCGTxC", b"TTACGTACTT"))[0][0] == base  # a, x -> 'A'
    # N matches every base, on either side: never less likely than the concrete base it replaces,
    # and exactly as likely as a read whose base agrees with every haplotype column it can face
    assert oracle.port_pairhmm(mk(b"ACNTAC", b"TTACGTACTT"))[0][0] >= base
    assert oracle.port_pairhmm(mk(b"ACGTAC", b"TTACNTACTT"))[0][0] >= base
    assert oracle.port_pairhmm(mk(b"N", b"ACGT"))[0][0] == oracle.port_pairhmm(mk(b"A", b"AAAA"))[0][0]
    assert oracle.port_pairhmm(mk(b"C", b"NNNN"))[0][0] == oracle.port_pairhmm(mk(b"C", b"CCCC"))[0][0]
