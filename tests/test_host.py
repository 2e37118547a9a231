"""CPU: host-side logic -- batch layout, fixtures, synthetic configurations, C-ABI surface."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import oracle
from gkl_b200 import fixtures, native, synth
from gkl_b200.batch import PairHmmBatch

ROOT = Path(__file__).resolve().parents[1]


def test_batch_from_lists_and_slices():
    b = PairHmmBatch.from_lists([b"ACGT", b"AC"], [b"\x1e" * 4, b"\x1e" * 2], [b"\x28" * 4, b"\x28" * 2],
                                [b"\x28" * 4, b"\x28" * 2], [b"\x0a" * 4, b"\x0a" * 2], [b"ACGTT", b"ACG", b"A"])
    b.validate()
    assert (b.n_reads, b.n_haps) == (2, 3)
    assert b.cells() == (4 + 2) * (5 + 3 + 1)
    s = b.read_slice(1, 2)
    s.validate()
    assert s.n_reads == 1 and bytes(s.read_bases) == b"AC" and s.n_haps == 3
    with pytest.raises(ValueError):
        PairHmmBatch.from_lists([b"ACGT"], [b"\x1e" * 3], [b"\x28" * 4], [b"\x28" * 4], [b"\x0a" * 4], [b"A"])


def test_golden_file_decoding_matches_gkl_test_harness(golden_pairhmm):
    # PairHmmUnitTest.java:206-212: quals - 33, read quals clamped to >= 6
    batches, expected = golden_pairhmm
    assert len(batches) == len(expected) == 104
    for b in batches:
        assert b.read_quals.min() >= 6
        assert b.ins_gop.max() < 94
    assert -15 < expected.min() < expected.max() < -1


def test_synthetic_configs_are_deterministic_and_shaped():
    c1 = synth.config1()
    assert (c1.n_reads, c1.n_haps, int(c1.read_lens[0]), int(c1.hap_lens[0])) == (1, 1, 50, 100)
    a, b = synth.config2(50, 16), synth.config2(50, 16)
    assert np.array_equal(a.read_bases, b.read_bases) and np.array_equal(a.hap_bases, b.hap_bases)
    assert set(a.read_lens) == {101} and a.hap_lens.min() >= 200 and a.hap_lens.max() <= 400
    regs = synth.config3(2)
    assert all(35 <= r.read_lens.min() and r.read_lens.max() <= 250 for r in regs)


def test_library_exports_every_symbol_of_the_header():
    header = (ROOT / "include" / "gklb_pairhmm.h").read_text()
    declared = set(re.findall(r"GKLB_API\s+[\w\s\*]+?\b(gklb_\w+)\s*\(", header))
    assert declared == set(native.EXPORTS), declared ^ set(native.EXPORTS)
    lib = native.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.gklb_version()
    # the PDHMM surface (include/gklb_pdhmm.h): exported by libgkl_pairhmm.so (which holds everything) and by
    # libgkl_pdhmm.so (the PDHMM engine + its JNI exports alone, under the name GKL's loader asks for)
    from gkl_b200 import pdhmm
    pd_header = (ROOT / "include" / "gklb_pdhmm.h").read_text()
    pd_declared = set(re.findall(r"GKLB_API\s+[\w\s\*]+?\b(gklb_\w+)\s*\(", pd_header))
    assert pd_declared == set(pdhmm.PD_EXPORTS), pd_declared ^ set(pdhmm.PD_EXPORTS)
    import ctypes
    pd_lib = ctypes.CDLL(str(native.LIB_PATH.with_name("libgkl_pdhmm.so")))
    common = {"gklb_last_error", "gklb_version", "gklb_device_count"}
    for name in pd_declared | common:
        assert hasattr(pd_lib, name), name
        assert hasattr(lib, name), name
    for name in ("initNative", "computeLikelihoodsNative", "computePDHMMNative", "doneNative"):
        assert hasattr(pd_lib, "Java_com_intel_gkl_pdhmm_IntelPDHMM_" + name)
    assert hasattr(pd_lib, "JNI_OnLoad") and not hasattr(pd_lib, "gklb_pairhmm_compute")
    # the Smith-Waterman surface (include/gklb_sw.h), third library name
    from gkl_b200 import smithwaterman
    sw_header = (ROOT / "include" / "gklb_sw.h").read_text()
    sw_declared = set(re.findall(r"GKLB_API\s+[\w\s\*]+?\b(gklb_\w+)\s*\(", sw_header))
    assert sw_declared == set(smithwaterman.SW_EXPORTS), sw_declared ^ set(smithwaterman.SW_EXPORTS)
    sw_lib = ctypes.CDLL(str(native.LIB_PATH.with_name("libgkl_smithwaterman.so")))
    for name in sw_declared | common:
        assert hasattr(sw_lib, name), name
        assert hasattr(lib, name), name
    assert hasattr(sw_lib, "Java_com_intel_gkl_smithwaterman_IntelSmithWaterman_alignNative") and hasattr(sw_lib, "JNI_OnLoad")


def test_host_tables_are_bit_identical_to_the_oracle():
    # Context.h:65-89,133-189 evaluated by two independent restatements
    ours, ref = native.tables(), oracle.port_tables()
    for k in ours:
        assert np.array_equal(ours[k], ref[k][:len(ours[k])]), k


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(native.GklbError) as ei:
        native.Engine(0)
    assert ei.value.code == native.ERR_NO_DEVICE
    rc = native.lib().gklb_pairhmm_compute(None, None)
    assert rc == native.ERR_STATE  # not initialised; never computes on the host


def test_product_package_does_not_touch_the_oracle():
    for p in (ROOT / "gkl_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cc", ".h") and p.is_file():
            text = p.read_text()
            assert "import oracle" not in text and "from oracle" not in text, p
            assert "libgklb_oracle" not in text and "libgkl_ref" not in text, p


def test_smithwaterman_mirror_validates_like_the_java_wrapper_and_refuses_without_a_gpu():
    # IntelSmithWaterman.java:122-145; no device call is made for rejected arguments
    import pytest
    from gkl_b200.pairhmm import IllegalArgumentException, NullPointerException
    from gkl_b200.smithwaterman import IntelSmithWaterman, SWOverhangStrategy, SWParameters
    sw = IntelSmithWaterman()
    p = SWParameters(10, -5, -10, -10)
    for args, exc in (((None, b"A", p, SWOverhangStrategy.IGNORE), NullPointerException),
                      ((b"A", None, p, SWOverhangStrategy.IGNORE), NullPointerException),
                      ((b"A", b"A", None, SWOverhangStrategy.IGNORE), NullPointerException),
                      ((b"A", b"A", p, None), NullPointerException),
                      ((b"", b"A", p, SWOverhangStrategy.IGNORE), IllegalArgumentException),
                      ((b"A" * 32768, b"A", p, SWOverhangStrategy.IGNORE), IllegalArgumentException),
                      ((b"A", b"A", SWParameters(65537, 0, 0, 0), SWOverhangStrategy.IGNORE), IllegalArgumentException),
                      ((b"A", b"A", p, 13), IllegalArgumentException)):
        with pytest.raises(exc):
            sw.align(*args)
    assert [s.value for s in SWOverhangStrategy] == [9, 10, 11, 12]  # getStrategy (:153-170)
    import torch
    if not torch.cuda.is_available():
        assert sw.load() is False  # no CPU path: GATK keeps its Java aligner
