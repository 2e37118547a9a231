"""CPU: pin the PDHMM oracle against the reference's golden files and its own compiled code."""
import numpy as np
import pytest

import oracle
from gkl_b200 import pdhmm_batch as pb
from tests.conftest import GOLDEN

TOL = 1e-4  # IntelPDHMMUnitTest.java:33 (absolute, log10 likelihood)
FILES = ["pdhmm_syn_990_1_2.txt", "pdhmm_syn_199_68_51.txt", "pdhmm_syn_1412_129_223.txt.gz"]
needs_ref = pytest.mark.skipif(not oracle.ref_available() and not oracle.REFERENCE_ROOT.is_dir(),
                               reason="oracle/_ref not built and /root/reference absent")


@pytest.mark.parametrize("name", FILES)
def test_port_matches_golden_pair_files(name):
    b, expected = pb.load_pdhmm_pairs_file(GOLDEN / name)
    got, status, _ = oracle.port_pdhmm(b, True, threads=4)
    assert status == 0
    assert np.abs(got - expected).max() <= TOL


def test_port_matches_pdhmm_new_cross_product_in_read_major_order():
    reads, haps, expected = pb.load_pdhmm_new(GOLDEN / "pdhmm_new.txt")
    assert (len(reads), len(haps), len(expected)) == (276, 48, 276 * 48)
    b = pb.PdhmmBatch.cross(reads, haps)
    got, status, _ = oracle.port_pdhmm(b, True, threads=4)
    assert status == 0 and np.abs(got - expected).max() <= TOL  # pins pair index r * H + h
    hap_major = got.reshape(276, 48).T.ravel()
    assert np.abs(hap_major - expected).max() > 1.0


@needs_ref
@pytest.mark.parametrize("name", FILES)
def test_port_is_the_reference_serial_path_bit_for_bit(name):
    b, _ = pb.load_pdhmm_pairs_file(GOLDEN / name)
    serial, rc, _ = oracle.ref_pdhmm(b, level=1)
    assert rc == 0
    assert np.array_equal(oracle.port_pdhmm(b, True)[0], serial)


@needs_ref
def test_row_reset_variant_is_the_reference_avx_path_and_where_they_differ():
    b, expected = pb.load_pdhmm_pairs_file(GOLDEN / "pdhmm_syn_1412_129_223.txt.gz")
    carry = oracle.port_pdhmm(b, True, threads=4)[0]
    reset = oracle.port_pdhmm(b, False, threads=4)[0]
    avx = oracle.ref_pdhmm(b, level=0)[0]
    assert np.abs(reset - avx).max() <= 1e-9
    differ = np.flatnonzero(np.abs(carry - reset) > 1e-9)
    # the two reference paths disagree exactly when the last haplotype base opens a deletion it never closes
    for k in differ:
        last = b.hap_pdbases[k * b.max_hap + b.hap_lengths[k] - 1]
        assert (last & 2) and not (last & 4)
    assert 0 < len(differ) < 20 and np.abs(carry - reset).max() < 1e-4
    assert np.abs(carry - expected).max() <= np.abs(reset - expected).max()  # the goldens follow the serial path
