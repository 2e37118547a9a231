"""CPU: libgkl_utils.so (gkl_b200/csrc/jni_utils.cc), the loader's prerequisite library -- the checks of the
reference's IntelGKLUtilsUnitTest.java (FTZ get/set round trip, isAvx* smoke, thread count) plus agreement of the
AVX answers with GKL's own common/avx.h where oracle/_ref is built."""
import ctypes as C
import os
import subprocess

import pytest

import oracle
from gkl_b200 import native

LIB = native.LIB_PATH.with_name("libgkl_utils.so")
P = "Java_com_intel_gkl_IntelGKLUtils_"


@pytest.fixture(scope="module")
def lib():
    l = C.CDLL(str(LIB))
    for name, res, args in (("getFlushToZeroNative", C.c_uint8, [C.c_void_p, C.c_void_p]),
                            ("setFlushToZeroNative", None, [C.c_void_p, C.c_void_p, C.c_uint8]),
                            ("isAvxSupportedNative", C.c_uint8, [C.c_void_p, C.c_void_p]),
                            ("isAvx2SupportedNative", C.c_uint8, [C.c_void_p, C.c_void_p]),
                            ("isAvx512SupportedNative", C.c_uint8, [C.c_void_p, C.c_void_p]),
                            ("getAvailableOmpThreadsNative", C.c_int32, [C.c_void_p, C.c_void_p])):
        f = getattr(l, P + name)
        f.restype, f.argtypes = res, args
    return l


def test_exports_exactly_the_six_symbols_intelgklutils_binds():
    out = subprocess.run(["nm", "-D", "--defined-only", str(LIB)], capture_output=True, text=True, check=True).stdout
    syms = {l.split()[-1] for l in out.splitlines() if " T " in l and "Java_" in l}
    want = {P + n for n in ("getFlushToZeroNative", "setFlushToZeroNative", "isAvxSupportedNative", "isAvx2SupportedNative",
                            "isAvx512SupportedNative", "getAvailableOmpThreadsNative")}  # IntelGKLUtils.java:109-114
    assert syms == want
    # no CUDA, no OpenMP runtime: it must load on a host that has neither
    deps = subprocess.run(["ldd", str(LIB)], capture_output=True, text=True).stdout
    assert "cuda" not in deps.lower() and "gomp" not in deps


def test_flush_to_zero_round_trip(lib):
    # IntelGKLUtilsUnitTest.simpleTest: set true / read true, set false / read false; restore what was there
    before = getattr(lib, P + "getFlushToZeroNative")(None, None)
    try:
        getattr(lib, P + "setFlushToZeroNative")(None, None, 1)
        assert getattr(lib, P + "getFlushToZeroNative")(None, None) == 1
        getattr(lib, P + "setFlushToZeroNative")(None, None, 0)
        assert getattr(lib, P + "getFlushToZeroNative")(None, None) == 0
    finally:
        getattr(lib, P + "setFlushToZeroNative")(None, None, before)


def test_avx_answers_and_thread_count(lib):
    avx, avx2, avx512 = (getattr(lib, P + n)(None, None) for n in ("isAvxSupportedNative", "isAvx2SupportedNative",
                                                                    "isAvx512SupportedNative"))
    assert avx in (0, 1) and avx2 in (0, 1) and avx512 in (0, 1)
    assert avx2 <= avx and avx512 <= avx2
    flags = open("/proc/cpuinfo").read() if os.path.exists("/proc/cpuinfo") else ""
    if " avx " in flags:
        assert avx == 1
    if " avx2 " in flags:
        assert avx2 == 1
    if oracle.ref_available():  # GKL's own is_avx512_supported() (common/avx.h:104-132)
        assert bool(avx512) == oracle.ref_avx512_supported()
    n = getattr(lib, P + "getAvailableOmpThreadsNative")(None, None)
    assert n == len(os.sched_getaffinity(0))
