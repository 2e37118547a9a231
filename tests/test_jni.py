"""The JNI face of libgkl_pairhmm.so, driven by a fake JVM (no JDK in this image).
CPU: symbols, failover without a GPU, exception classes.  GPU: results through the JNI path."""
import subprocess

import numpy as np
import pytest

import oracle
from gkl_b200 import native, synth
from tests import jni_fake

LIB = native.LIB_PATH


def has_gpu():
    import torch
    return torch.cuda.is_available()


def test_exports_exactly_the_symbols_intelpairhmm_binds():
    # IntelPairHmm.java:157-164 declares initNative, computeLikelihoodsNative, doneNative
    out = subprocess.run(["nm", "-D", "--defined-only", str(LIB)], capture_output=True, text=True, check=True).stdout
    syms = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for s in ("Java_com_intel_gkl_pairhmm_IntelPairHmm_initNative",
              "Java_com_intel_gkl_pairhmm_IntelPairHmm_computeLikelihoodsNative",
              "Java_com_intel_gkl_pairhmm_IntelPairHmm_doneNative", "JNI_OnLoad"):
        assert s in syms
    # IntelPDHMM.java:206-216 (the same binary is installed as libgkl_pdhmm.so)
    for s in ("initNative", "computeLikelihoodsNative", "computePDHMMNative", "doneNative"):
        assert "Java_com_intel_gkl_pdhmm_IntelPDHMM_" + s in syms
    # IntelSmithWaterman.java:183-186 (the same binary is installed as libgkl_smithwaterman.so)
    for s in ("initNative", "alignNative", "doneNative"):
        assert "Java_com_intel_gkl_smithwaterman_IntelSmithWaterman_" + s in syms
    assert not [s for s in syms if s.startswith("Java_") and "IntelPairHmm" not in s and "IntelPDHMM" not in s
                and "IntelSmithWaterman" not in s]


def test_library_refuses_to_load_into_a_jvm_without_a_gpu():
    if has_gpu():
        pytest.skip("a GPU is present")
    assert jni_fake.onload(LIB) == -1  # JNI_ERR -> UnsatisfiedLinkError -> IntelPairHmm.load() == false


def test_missing_field_is_illegal_argument_exception():
    b = synth.config2(3, 2)
    rc, _, cls, msg, leaks = jni_fake.pairhmm(LIB, b, fault=1)
    assert rc == 1 and cls == "java/lang/IllegalArgumentException" and msg == "Unable to get field ID"  # JavaData.h:127-133
    assert leaks == (0, 0)


def test_compute_before_init_and_no_device_raise_java_exceptions():
    b = synth.config2(3, 2)
    rc, _, cls, _, leaks = jni_fake.pairhmm(LIB, b, fault=5)
    assert rc == 1 and cls == "java/lang/IllegalStateException" and leaks == (0, 0)
    if not has_gpu():
        rc, _, cls, msg, leaks = jni_fake.pairhmm(LIB, b)
        assert rc == 1 and cls == "java/lang/RuntimeException" and "CUDA" in msg and leaks == (0, 0)


@pytest.mark.gpu
def test_onload_accepts_with_a_gpu():
    assert jni_fake.onload(LIB) == 0x00010006


@pytest.mark.gpu
@pytest.mark.parametrize("use_double", [False, True])
def test_results_through_the_jni_path(use_double):
    b = synth.random_batch(61, 40, 12, low_quality=0.05, unrelated=0.2)
    rc, out, cls, msg, leaks = jni_fake.pairhmm(LIB, b, use_double)
    assert rc == 0, (cls, msg)
    assert leaks == (0, 0)  # every local reference deleted, the output array released (copy-back mode 0)
    ref = (oracle.ref_pairhmm if oracle.ref_available() else oracle.port_pairhmm)(b, use_double)[0]
    assert (np.abs(out - ref) / np.abs(ref)).max() <= 1e-5


@pytest.mark.gpu
def test_large_calls_are_pipelined_in_read_blocks(monkeypatch):
    # GKLB_JNI_BLOCK_READS=50: a 330-read call is marshalled and submitted in 7 blocks on two alternating engines
    monkeypatch.setenv("GKLB_JNI_BLOCK_READS", "50")
    b = synth.random_batch(62, 330, 9, read_len=(20, 140), hap_len=(60, 200), unrelated=0.1)
    rc, out, cls, msg, leaks = jni_fake.pairhmm(LIB, b)
    assert rc == 0 and leaks == (0, 0), (cls, msg)
    ref = (oracle.ref_pairhmm if oracle.ref_available() else oracle.port_pairhmm)(b, threads=oracle.host_threads())[0]
    assert (np.abs(out - ref) / np.abs(ref)).max() <= 1e-5
    rc, _, cls, _, leaks = jni_fake.pairhmm(LIB, b, fault=2)  # a null read in the middle of block 3
    assert rc == 1 and cls == "java/lang/NullPointerException" and leaks == (0, 0)


@pytest.mark.gpu
def test_pipelined_calls_run_on_the_configured_device(monkeypatch):
    """The read-block pipeline borrows its two engines from the pool, i.e. on the device the pool was configured
    with (GKLB_DEVICES), not on device 0."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    monkeypatch.setenv("GKLB_DEVICES", "1")
    monkeypatch.setenv("GKLB_JNI_BLOCK_READS", "40")
    b = synth.random_batch(63, 200, 7, read_len=(20, 140), hap_len=(60, 200))
    free0 = torch.cuda.mem_get_info(0)[0]
    rc, out, cls, msg, leaks = jni_fake.pairhmm(LIB, b)
    assert rc == 0 and leaks == (0, 0), (cls, msg)
    assert torch.cuda.mem_get_info(0)[0] >= free0 - (64 << 20)   # nothing was allocated on device 0
    ref = (oracle.ref_pairhmm if oracle.ref_available() else oracle.port_pairhmm)(b, threads=oracle.host_threads())[0]
    assert (np.abs(out - ref) / np.abs(ref)).max() <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("fault,cls", [(2, "java/lang/NullPointerException"), (3, "java/lang/OutOfMemoryError"),
                                       (4, "java/lang/IllegalArgumentException"), (6, "java/lang/IllegalArgumentException")])
def test_fault_injection_maps_to_gkl_exception_classes(fault, cls):
    b = synth.config2(5, 3)
    rc, _, got, _, leaks = jni_fake.pairhmm(LIB, b, fault=fault)
    assert rc == 1 and got == cls and leaks == (0, 0)


@pytest.mark.gpu
def test_done_is_idempotent_and_reinitialise_works():
    b = synth.config2(6, 4)
    rc, out, cls, msg, leaks = jni_fake.pairhmm(LIB, b, fault=7)
    assert rc == 0 and leaks == (0, 0), (cls, msg)
    assert (np.abs(out - oracle.port_pairhmm(b)[0]) / np.abs(out)).max() <= 1e-5


PD_LIB = native.LIB_PATH.with_name("libgkl_pdhmm.so")


def test_pdhmm_missing_field_is_illegal_argument_exception():
    from gkl_b200 import pdhmm_batch as pb
    from tests.conftest import GOLDEN
    b, _ = pb.load_pdhmm_pairs_file(GOLDEN / "pdhmm_syn_990_1_2.txt", limit=4)
    rc, _, cls, msg, leaks = jni_fake.pdhmm(PD_LIB, b, fault=1)
    assert rc == 1 and cls == "java/lang/IllegalArgumentException" and msg == "Unable to get field ID" and leaks == (0, 0)


@pytest.mark.gpu
def test_pdhmm_flat_and_object_api_through_jni():
    from gkl_b200 import pdhmm_batch as pb
    from tests.conftest import GOLDEN
    b, expected = pb.load_pdhmm_pairs_file(GOLDEN / "pdhmm_syn_199_68_51.txt")
    rc, out, cls, msg, leaks = jni_fake.pdhmm(PD_LIB, b)
    assert rc == 0 and leaks == (0, 0), (cls, msg)
    assert np.abs(out - expected).max() <= 1e-4  # IntelPDHMMUnitTest.java:33
    rc, _, cls, _, leaks = jni_fake.pdhmm(PD_LIB, b, fault=2)
    assert rc == 1 and cls == "java/lang/IllegalArgumentException" and leaks == (0, 0)
    reads, haps, exp = pb.load_pdhmm_new(GOLDEN / "pdhmm_new.txt")
    reads, haps = reads[:40], haps[:12]
    ops = pb.PdhmmBatch.from_pairs([(h[0], h[1], b"", b"", b"", b"", b"") for h in haps])
    rds = pb.PdhmmBatch.from_pairs([(b"", b"", *r) for r in reads])
    operands = pb.PdhmmBatch(ops.hap_bases, ops.hap_pdbases, rds.read_bases, rds.read_qual, rds.read_ins_qual,
                             rds.read_del_qual, rds.gcp, ops.hap_lengths, rds.read_lengths, ops.max_hap, rds.max_read)
    rc, out, cls, msg, leaks = jni_fake.pdhmm(PD_LIB, operands, object_api=True, n_reads=40, n_haps=12)
    assert rc == 0 and leaks == (0, 0), (cls, msg)
    want = exp.reshape(276, 48)[:40, :12].ravel()
    assert np.abs(out - want).max() <= 1e-4


def test_smithwaterman_without_a_gpu_raises_from_initnative():
    if has_gpu():
        pytest.skip("a GPU is present")
    rc, _, _, cls, msg, leaks = jni_fake.smithwaterman(LIB, b"ACGT", b"ACGT", (3, -1, -4, -3), 9)
    assert rc == 1 and cls == "java/lang/RuntimeException" and leaks == (0, 0)


@pytest.mark.gpu
def test_smithwaterman_align_through_jni():
    from tests.test_oracle_sw import load_sw_golden, PARAMS, STRATEGIES
    refs, alts, exp = load_sw_golden()
    for k in (0, 7, 100, len(refs) - 1):
        col = 0
        for p in PARAMS:
            for st in STRATEGIES:
                rc, cigar, off, cls, msg, leaks = jni_fake.smithwaterman(LIB, refs[k], alts[k], p, st)
                assert rc == 0, (cls, msg)
                assert (cigar.rstrip(b"\x00").decode(), off) == exp[k][col]
                assert leaks == (0, 0)
                col += 1
    # GetPrimitiveArrayCritical failing -> IllegalArgumentException("Arrays aren't valid."), IntelSmithWaterman.cc:82-104
    rc, _, _, cls, msg, leaks = jni_fake.smithwaterman(LIB, refs[0], alts[0], PARAMS[0], 9, fault=1)
    assert rc == 1 and cls == "java/lang/IllegalArgumentException" and msg == "Arrays aren't valid." and leaks == (0, 0)


@pytest.mark.gpu
def test_concurrent_callers_through_the_jni_layer(monkeypatch):
    """8 threads x 32 regions over the fake JVM, every thread its own JNIEnv and IntelPairHmm instance (init at the
    start, done at the end, while other threads are still computing): the engine pool must hand every call its own
    engine, results must be bit-identical to serial calls, no reference or pin may leak.  GKL's state is read-only
    after initNative, so GATK may and does call computeLikelihoods concurrently (IntelPairHmm.java:65)."""
    import torch
    if torch.cuda.device_count() >= 2:
        monkeypatch.setenv("GKLB_DEVICES", "all")   # concurrent calls spread over the GPUs of the box
    regions = synth.config3(32, seed=21)
    eng = native.Engine(0, False)
    serial = [eng.compute(r) for r in regions]
    eng.close()
    failed, outs, msg, leaks = jni_fake.pairhmm_threads(LIB, regions, n_threads=8, rounds=2)
    assert failed == 0, msg
    assert leaks == (0, 0)
    for a, b in zip(serial, outs):
        assert np.array_equal(a, b)
    assert native.lib().gklb_pairhmm_engines_alive() == 0   # the last doneNative freed the pool
