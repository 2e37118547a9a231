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
    assert st.pairs == 276 * 48 and st.kernel_launches in (1, 2)  # k_pdhmm3 (+ k_pdhmm2 for deferred haplotypes)


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


def _truncated_pairs(b, max_rows):
    """The pairs of a flat batch with every read cut to its first 60..max_rows rows."""
    pairs = []
    for k in range(b.n):
        hl, rl = int(b.hap_lengths[k]), min(int(b.read_lengths[k]), 60 + k % (max_rows - 59))
        h0, r0 = k * b.max_hap, k * b.max_read
        pairs.append((b.hap_bases[h0:h0 + hl], b.hap_pdbases[h0:h0 + hl], b.read_bases[r0:r0 + rl],
                      b.read_qual[r0:r0 + rl], b.read_ins_qual[r0:r0 + rl], b.read_del_qual[r0:r0 + rl],
                      b.gcp[r0:r0 + rl]))
    return pb.PdhmmBatch.from_pairs(pairs)


@pytest.mark.parametrize("max_rows", [124, 155, 186])
@pytest.mark.parametrize("row_state", ["carry", "reset"])
def test_single_pass_kernel_on_the_hardest_golden_pairs(row_state, max_rows):
    """Reads of at most 124 / 155 / 186 rows take the haplotype-major kernel (k_pdhmm2 with 4 / 5 / 6 rows per lane).  The synthetic golden file has random PD
    bytes (nested / unclosed spans, negative bytes), lower-case bases and quals up to 93, but only reads of 223+ rows:
    cut its reads to 60..max_rows rows and run them through that kernel in both row-state modes, and through the
    pair-at-a-time kernel, and compare everything with the restatement of the reference's scalar path."""
    full, _ = pb.load_pdhmm_pairs_file(GOLDEN / "pdhmm_syn_1412_129_223.txt.gz")
    b = _truncated_pairs(full, max_rows)
    assert b.max_read == max_rows and b.n == full.n
    ref = oracle.port_pdhmm(b, row_state == "carry", threads=oracle.host_threads())[0]
    got = {}
    for kernel in ("2", "1"):
        os.environ["GKLB_PDHMM_ROW_STATE"] = row_state
        os.environ["GKLB_PDHMM_KERNEL"] = kernel
        try:
            h = IntelPDHMM()
            h.initialize(None)
            got[kernel] = h.compute_batch(b)
            h.done()
        finally:
            os.environ.pop("GKLB_PDHMM_ROW_STATE")
            os.environ.pop("GKLB_PDHMM_KERNEL")
    assert np.abs(got["2"] - ref).max() <= 1e-9
    assert np.abs(got["1"] - ref).max() <= 1e-9
    assert np.array_equal(got["1"], got["2"])  # same arithmetic, cell for cell


def test_single_pass_kernel_cross_layout_blocks():
    """Cross layout (computeLikelihoods): a task of k_pdhmm2 is one haplotype x a block of reads; ragged read counts,
    one-column haplotypes and PD spans at the haplotype ends must land at r * H + h."""
    from gkl_b200 import synth
    reads, haps = synth.config5(77, 19, seed=11)
    rng = np.random.default_rng(5)
    # shorten some reads / haplotypes, put spans at the very ends
    cut = [int(rng.integers(1, len(r[0]) + 1)) if i % 5 == 0 else len(r[0]) for i, r in enumerate(reads)]
    reads = [tuple(x[:n] for x in r) for r, n in zip(reads, cut)]
    haps = list(haps)
    haps[0] = (haps[0][0][:1], np.array([6], dtype=np.int8))
    haps[1] = (haps[1][0][:2], np.array([2, 4], dtype=np.int8))
    hb, pdb = haps[2][0].copy(), haps[2][1].copy()
    pdb[-1] = 2   # a deletion that opens on the last column and never closes: the row-carry corner (P3)
    haps[2] = (hb, pdb)
    hb, pdb = haps[3][0].copy(), haps[3][1].copy()
    pdb[0] = 2
    pdb[3] = 4
    haps[3] = (hb, pdb)
    flat = pb.PdhmmBatch.cross(reads, haps)
    ref = oracle.port_pdhmm(flat, True, threads=oracle.host_threads())[0]
    rd = [PDReadDataHolder(*(x.tobytes() for x in r)) for r in reads]
    hp = [PDHaplotypeDataHolder(h[0].tobytes(), h[1].tobytes()) for h in haps]
    h = IntelPDHMM()
    h.initialize(None)
    out = np.zeros(len(rd) * len(hp))
    h.computeLikelihoods(rd, hp, out)
    flat_out = h.compute_batch(flat)
    h.done()
    assert np.abs(out - ref).max() <= 1e-9
    assert np.abs(flat_out - ref).max() <= 1e-9


def _corner_batch(odd_read_bytes: bool, snp_columns: bool):
    from gkl_b200 import synth
    reads, haps = synth.config5(61, 24, seed=17)
    rng = np.random.default_rng(18)
    reads = [tuple(x.copy() for x in r) for r in reads]
    haps = [(h[0].copy(), h[1].copy()) for h in haps]
    if not snp_columns:
        for _, pdb in haps:
            pdb &= 6   # keep the deletion spans only
    for i, r in enumerate(reads):
        n = len(r[0])
        if i % 4 == 0:   # lower case and N in the read; optionally bytes outside ACGTacgtN
            pos = rng.integers(0, n, size=6)
            r[0][pos[:2]] = ord("a") if i % 8 else ord("t")
            r[0][pos[2:4]] = ord("N")
            if odd_read_bytes:
                r[0][pos[4:]] = np.array([ord("X"), 0 if i % 8 else ord("n")], dtype=np.int8)
        if i % 7 == 0:
            reads[i] = tuple(x[:int(rng.integers(1, n))] for x in r)
    for hidx in (0, 1, 2):   # many kinds of columns: every allele combination on every base, other bytes, N
        hb, pdb = haps[hidx]
        pos = rng.choice(len(hb), size=40, replace=False)
        if snp_columns:
            pdb[pos[:30]] = (1 | (rng.integers(1, 16, size=30) << 3)).astype(np.int8)
        hb[pos[30:34]] = ord("N")
        hb[pos[34:37]] = ord("X")
        hb[pos[37:]] = np.array([ord("c"), ord("g"), 0], dtype=np.int8)
    haps[3][1][-1] |= 2            # opens on the last column: rows start INSIDE
    haps[4][1][-1] |= 4            # closes on the last column: rows start AFTER
    haps[5][1][-3] |= 2
    haps[6][1][0] |= 2; haps[6][1][0] |= 4; haps[6][1][1] |= 2; haps[6][1][2] |= 4   # adjacent spans at the start
    haps[7][1][10] |= 2; haps[7][1][14] |= 2; haps[7][1][20] |= 4; haps[7][1][21] |= 4   # nested start, double end
    haps[8] = (haps[8][0][:1], np.array([0], dtype=np.int8))
    haps[9] = (haps[9][0][:17], haps[9][1][:17])
    return reads, haps


@pytest.mark.parametrize("odd_read_bytes,snp_columns", [(False, True), (True, False)])
def test_two_reads_per_warp_kernel_corners(odd_read_bytes, snp_columns):
    """k_pdhmm3 (cross layout, reads of at most 105 rows): odd read counts, haplotypes that end inside / right after a
    deletion (deferred to k_pdhmm2 by the host), haplotypes with more kinds of columns than the prior table holds
    (deferred by the kernel), lower case and N on both sides, adjacent and nested spans; and, without SNP columns (where
    the reference would reject them), read bytes outside ACGTacgtN meeting identical haplotype bytes.  Against the
    restatement of the reference's scalar path and against k_pdhmm2 alone."""
    reads, haps = _corner_batch(odd_read_bytes, snp_columns)
    flat = pb.PdhmmBatch.cross(reads, haps)
    rd = [PDReadDataHolder(*(x.tobytes() for x in r)) for r in reads]
    hp = [PDHaplotypeDataHolder(h[0].tobytes(), h[1].tobytes()) for h in haps]
    for row_state in ("carry", "reset"):
        ref, rc, _ = oracle.port_pdhmm(flat, row_state == "carry", threads=oracle.host_threads())
        assert rc == 0
        got = {}
        for kernel in ("3", "2"):
            os.environ["GKLB_PDHMM_ROW_STATE"] = row_state
            os.environ["GKLB_PDHMM_KERNEL"] = kernel
            try:
                h = IntelPDHMM()
                h.initialize(None)
                out = np.zeros(len(rd) * len(hp))
                h.computeLikelihoods(rd, hp, out)
                got[kernel] = out
                launches = h.stats().kernel_launches
                h.done()
            finally:
                os.environ.pop("GKLB_PDHMM_ROW_STATE")
                os.environ.pop("GKLB_PDHMM_KERNEL")
            assert launches == (2 if kernel == "3" else 1)
        assert np.abs(got["3"] - ref).max() <= 1e-9
        assert np.abs(got["3"] - got["2"]).max() <= 1e-9   # k_pdhmm3 folds the deletion state: not bit-identical


def test_unexpected_read_base_at_an_snp_column_is_rejected_like_the_reference(hmm):
    """pdhmm-serial.cc:228-252: a read byte outside ACGTacgt (and not N, and not equal to the haplotype byte) on a column
    with SNP alleles is PDHMM_INPUT_DATA_ERROR -> IllegalArgumentException (IntelPDHMM.cc:102-105,222-226), for both
    entry points; the same bytes without such a column are fine."""
    reads, haps = _corner_batch(True, True)
    flat = pb.PdhmmBatch.cross(reads, haps)
    assert oracle.port_pdhmm(flat, True, threads=oracle.host_threads())[1] == 2   # PDHMM_INPUT_DATA_ERROR
    rd = [PDReadDataHolder(*(x.tobytes() for x in r)) for r in reads]
    hp = [PDHaplotypeDataHolder(h[0].tobytes(), h[1].tobytes()) for h in haps]
    with pytest.raises(IllegalArgumentException):
        hmm.computeLikelihoods(rd, hp, np.zeros(len(rd) * len(hp)))
    with pytest.raises(IllegalArgumentException):
        hmm.compute_batch(flat)
    # one pair of the flat layout whose odd byte only meets an identical haplotype byte on the SNP column: accepted
    one = pb.PdhmmBatch.from_pairs([(np.frombuffer(b"ACXGT", dtype=np.int8), np.array([0, 0, 33, 0, 0], dtype=np.int8),
                                     np.frombuffer(b"ACXGT", dtype=np.int8), *(np.full(5, q, dtype=np.int8) for q in (30, 40, 40, 10)))])
    ref, rc, _ = oracle.port_pdhmm(one, True, threads=1)
    assert rc == 0 and np.abs(hmm.compute_batch(one) - ref).max() <= 1e-9


@pytest.mark.parametrize("seed", range(16))
def test_fuzz_cross_batches_against_the_scalar_restatement(seed):
    """Random cross-layout batches for k_pdhmm3 (both instantiations) and its deferrals: random PD bytes (SNP alleles
    in every combination, nested / unclosed / adjacent deletion spans, spans at both ends), N and lower case on both
    sides, ragged lengths from 1, qualities over a wide range; both row-state modes."""
    rng = np.random.default_rng(7000 + seed)
    max_read = 105 if seed % 2 == 0 else 155
    n_reads, n_haps = int(rng.integers(1, 48)), int(rng.integers(1, 14))
    letters = np.frombuffer(b"ACGT", dtype=np.int8)
    haps = []
    for _ in range(n_haps):
        L = int(rng.integers(1, 320))
        hb = letters[rng.integers(0, 4, size=L)].copy()
        u = rng.random(L)
        hb[u < 0.02] = ord("N")
        low = (u > 0.02) & (u < 0.05)
        hb[low] |= 0x20
        pdb = np.zeros(L, dtype=np.int8)
        v = rng.random(L)
        snp = v < 0.06
        pdb[snp] = (1 | (rng.integers(1, 16, size=int(snp.sum())) << 3)).astype(np.int8)
        pdb[(v > 0.06) & (v < 0.085)] |= 2
        pdb[(v > 0.085) & (v < 0.11)] |= 4
        pdb[(v > 0.11) & (v < 0.115)] |= 6
        if rng.random() < 0.3:
            pdb[-1] |= int(rng.choice([2, 4, 6]))
        if rng.random() < 0.3:
            pdb[0] |= int(rng.choice([2, 4, 6]))
        haps.append((hb, pdb))
    reads = []
    for _ in range(n_reads):
        R = int(rng.integers(1, max_read + 1))
        hb = haps[int(rng.integers(0, n_haps))][0]
        s = int(rng.integers(0, max(1, len(hb) - R + 1)))
        rb = np.resize(hb[s:s + R].copy() if len(hb[s:s + R]) else letters, R).astype(np.int8)
        rb[rb == ord("n")] = ord("N")
        mut = rng.random(R) < 0.05
        rb[mut] = letters[rng.integers(0, 4, size=int(mut.sum()))]
        rb[rng.random(R) < 0.02] = ord("N")
        q = lambda lo, hi: rng.integers(lo, hi, size=R).astype(np.int8)
        # insertion qualities within 30 dB let the engine pick the kernel variant with the folded insertion state
        ins = q(18, 46) if seed % 4 < 2 else q(5, 61)
        reads.append((rb, q(2, 61), ins, q(5, 61), q(1, 30)))
    flat = pb.PdhmmBatch.cross(reads, haps)
    rd = [PDReadDataHolder(*(x.tobytes() for x in r)) for r in reads]
    hp = [PDHaplotypeDataHolder(h[0].tobytes(), h[1].tobytes()) for h in haps]
    for row_state in ("carry", "reset"):
        ref, rc, _ = oracle.port_pdhmm(flat, row_state == "carry", threads=oracle.host_threads())
        assert rc == 0
        os.environ["GKLB_PDHMM_ROW_STATE"] = row_state
        try:
            h = IntelPDHMM()
            h.initialize(None)
            out = np.zeros(len(rd) * len(hp))
            h.computeLikelihoods(rd, hp, out)
            name = h.kernel_name()
            h.done()
        finally:
            os.environ.pop("GKLB_PDHMM_ROW_STATE")
        assert name.startswith("k_pdhmm3<16,7" if max_read == 105 and max(len(r[0]) for r in reads) <= 105 else "k_pdhmm3")
        assert name.endswith(",wfold>") == (seed % 4 < 2)
        ok = np.isfinite(ref)
        assert np.array_equal(np.isfinite(out), ok)
        assert np.abs(out[ok] - ref[ok]).max() <= 1e-9, (seed, row_state)


def test_folded_insertion_state_variant_against_the_unfolded_one():
    """k_pdhmm3<..., wfold> (insertion qualities within 30 dB, the config-5 shape) against the variant without the fold
    (GKLB_PDHMM_WFOLD=0) and the scalar restatement; a read whose insertion qualities span more switches the fold off
    for the batch."""
    from gkl_b200 import synth
    reads, haps = synth.config5(90, 20, seed=23)
    flat = pb.PdhmmBatch.cross(reads, haps)
    ref, rc, _ = oracle.port_pdhmm(flat, True, threads=oracle.host_threads())
    assert rc == 0
    rd = [PDReadDataHolder(*(x.tobytes() for x in r)) for r in reads]
    hp = [PDHaplotypeDataHolder(h[0].tobytes(), h[1].tobytes()) for h in haps]
    got, names = {}, {}
    for mode in ("1", "0"):
        os.environ["GKLB_PDHMM_WFOLD"] = mode
        try:
            h = IntelPDHMM()
            h.initialize(None)
            out = np.zeros(len(rd) * len(hp))
            h.computeLikelihoods(rd, hp, out)
            got[mode], names[mode] = out, h.kernel_name()
            h.done()
        finally:
            os.environ.pop("GKLB_PDHMM_WFOLD")
    assert names["1"].endswith(",wfold>") and not names["0"].endswith(",wfold>")
    assert np.abs(got["1"] - ref).max() <= 1e-9 and np.abs(got["0"] - ref).max() <= 1e-9
    assert np.abs(got["1"] - got["0"]).max() <= 1e-9
    wide = list(reads)
    r0 = tuple(x.copy() for x in wide[0])
    r0[2][0], r0[2][1] = 5, 60   # 55 dB between two insertion qualities of one read
    wide[0] = r0
    h = IntelPDHMM()
    h.initialize(None)
    out = np.zeros(len(rd) * len(hp))
    h.computeLikelihoods([PDReadDataHolder(*(x.tobytes() for x in r)) for r in wide], hp, out)
    name = h.kernel_name()
    h.done()
    assert not name.endswith(",wfold>")
    ref2 = oracle.port_pdhmm(pb.PdhmmBatch.cross(wide, haps), True, threads=oracle.host_threads())[0]
    assert np.abs(out - ref2).max() <= 1e-9


def test_folded_insertion_state_at_the_top_of_the_fp64_range():
    """The fold's scale argument at its limits: insertion qualities alternating over exactly 30 dB, gap continuation
    penalties of 1 and 2 (long geometric sums of the folded states), haplotypes of 1..8 columns (the initial condition 2^1020 / H
    at its largest) and full-length reads -- no overflow, same values as the scalar restatement."""
    rng = np.random.default_rng(77)
    letters = np.frombuffer(b"ACGT", dtype=np.int8)
    haps = []
    for L in (1, 1, 2, 3, 5, 8, 40, 300):
        hb = letters[rng.integers(0, 4, size=L)].copy()
        pdb = np.zeros(L, dtype=np.int8)
        if L >= 3:
            pdb[1] = 33
        haps.append((hb, pdb))
    reads = []
    for k in range(24):
        R = 105 if k % 3 else int(rng.integers(1, 106))
        rb = letters[rng.integers(0, 4, size=R)].copy()
        lo, hi = (15, 45) if k % 2 else (0, 30)
        ins = np.where(np.arange(R) % 2 == (k // 2) % 2, lo, hi).astype(np.int8)
        gcp = np.full(R, 0 if k % 8 == 7 else 1 + (k % 3 == 0), dtype=np.int8)   # 0: no way back from an insertion
        reads.append((rb, rng.integers(2, 61, size=R).astype(np.int8), ins, rng.integers(0, 61, size=R).astype(np.int8), gcp))
    flat = pb.PdhmmBatch.cross(reads, haps)
    ref, rc, _ = oracle.port_pdhmm(flat, True, threads=oracle.host_threads())
    assert rc == 0 and np.isfinite(ref).mean() > 0.8
    h = IntelPDHMM()
    h.initialize(None)
    out = np.zeros(len(reads) * len(haps))
    h.computeLikelihoods([PDReadDataHolder(*(x.tobytes() for x in r)) for r in reads],
                         [PDHaplotypeDataHolder(x[0].tobytes(), x[1].tobytes()) for x in haps], out)
    name = h.kernel_name()
    h.done()
    assert name.endswith(",wfold>")
    ok = np.isfinite(ref)
    assert np.array_equal(np.isfinite(out), ok) and np.abs(out[ok] - ref[ok]).max() <= 1e-9
