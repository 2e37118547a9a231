"""GPU: the PDHMM CUDA path through the C-ABI against the reference's golden files (1e-4 absolute,
IntelPDHMMUnitTest.java:33) and, much tighter, against the bit-exact restatement of its serial path."""
import os

import numpy as np
import pytest

import oracle
from gkl_b200 import pdhmm_batch as pb
from gkl_b200.pairhmm import IllegalArgumentException, NullPointerException
from gkl_b200.pdhmm import IntelPDHMM, PDHaplotypeDataHolder, PDHMMNativeArguments, PDReadDataHolder
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
FILES = ["pdhmm_syn_990_1_2.txt", "pdhmm_syn_199_68_51.txt", "pdhmm_syn_1412_129_223.txt.gz"]


@pytest.fixture(scope="module")
def hmm():
    h = IntelPDHMM()
    assert h.load()
    h.initialize(PDHMMNativeArguments(2, 0, 0, 10))  # IntelPDHMMUnitTest.java:91-102
    yield h
    h.done()


@pytest.mark.parametrize("name", FILES)
def test_golden_pair_files_flat_api(hmm, name):
    b, expected = pb.load_pdhmm_pairs_file(GOLDEN / name)
    got = hmm.computePDHMM(b.hap_bases, b.hap_pdbases, b.read_bases, b.read_qual, b.read_ins_qual, b.read_del_qual,
                           b.gcp, b.hap_lengths, b.read_lengths, b.n, b.max_hap, b.max_read)
    assert np.abs(got - expected).max() <= 1e-4
    serial = oracle.port_pdhmm(b, True, threads=oracle.host_threads())[0]  # == GKL's scalar path bit for bit
    assert np.abs(got - serial).max() <= 1e-9  # includes the reads of up to 442 rows (multi-pass)


def test_pdhmm_new_object_api_cross_product(hmm):
    reads, haps, expected = pb.load_pdhmm_new(GOLDEN / "pdhmm_new.txt")
    rd = [PDReadDataHolder(*(bytes(x.astype(np.uint8)) for x in r)) for r in reads]
    hp = [PDHaplotypeDataHolder(bytes(h[0].astype(np.uint8)), bytes(h[1].astype(np.uint8))) for h in haps]
    out = np.zeros(len(rd) * len(hp))
    hmm.computeLikelihoods(rd, hp, out)
    assert np.abs(out - expected).max() <= 1e-4  # r * H + h order
    st = hmm.stats()
    assert st.pairs == 276 * 48 and st.kernel_launches == 1


def test_argument_validation_matches_the_java_wrapper(hmm):
    b, _ = pb.load_pdhmm_pairs_file(GOLDEN / "pdhmm_syn_990_1_2.txt", limit=8)
    args = [b.hap_bases, b.hap_pdbases, b.read_bases, b.read_qual, b.read_ins_qual, b.read_del_qual, b.gcp,
            b.hap_lengths, b.read_lengths]
    with pytest.raises(NullPointerException):
        hmm.computePDHMM(None, *args[1:], b.n, b.max_hap, b.max_read)
    with pytest.raises(IllegalArgumentException):
        hmm.computePDHMM(b.hap_bases[:-1], *args[1:], b.n, b.max_hap, b.max_read)
    with pytest.raises(IllegalArgumentException):
        hmm.computePDHMM(*[a[:0] for a in args], 0, b.max_hap, b.max_read)
    bad = b.read_ins_qual.copy()
    bad[0] = -3  # negative quality: PDHMM_INPUT_DATA_ERROR -> IllegalArgumentException (pdhmm-serial.cc:184-198)
    with pytest.raises(IllegalArgumentException):
        hmm.computePDHMM(b.hap_bases, b.hap_pdbases, b.read_bases, b.read_qual, bad, b.read_del_qual, b.gcp,
                         b.hap_lengths, b.read_lengths, b.n, b.max_hap, b.max_read)
    ok = hmm.computePDHMM(*args, b.n, b.max_hap, b.max_read)
    assert np.all(np.isfinite(ok))


def test_row_reset_semantics_switch():
    """GKLB_PDHMM_ROW_STATE=reset reproduces the reference's AVX paths where they differ from its scalar path."""
    b, _ = pb.load_pdhmm_pairs_file(GOLDEN / "pdhmm_syn_1412_129_223.txt.gz")
    os.environ["GKLB_PDHMM_ROW_STATE"] = "reset"
    try:
        h = IntelPDHMM()
        h.initialize(None)
        got = h.compute_batch(b)
        h.done()
    finally:
        os.environ.pop("GKLB_PDHMM_ROW_STATE")
    reset = oracle.port_pdhmm(b, False, threads=oracle.host_threads())[0]
    assert np.abs(got - reset).max() <= 1e-9
