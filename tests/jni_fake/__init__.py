"""TEST INFRASTRUCTURE: ctypes driver of the fake JVM (tests/jni_fake/fake_jvm.cc)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SO = HERE / "libfakejvm.so"


def build():
    src = HERE / "fake_jvm.cc"
    if not SO.exists() or SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-o", str(SO), str(src), "-ldl"], check=True)


def _lib():
    build()
    l = C.CDLL(str(SO))
    l.fakejvm_onload.argtypes = [C.c_char_p]
    return l


def onload(lib_path) -> int:
    return _lib().fakejvm_onload(str(lib_path).encode())


def pairhmm(lib_path, b, use_double=False, fault=0):
    """Returns (rc, out, exception_class, exception_message, (leaked_local_refs, leaked_pins))."""
    l = _lib()
    out = np.zeros(max(1, b.n_reads * b.n_haps), dtype=np.float64)
    ec, em = C.create_string_buffer(256), C.create_string_buffer(256)
    leaks = (C.c_long * 2)()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = l.fakejvm_pairhmm(str(lib_path).encode(), b.n_reads, b.n_haps, p(b.read_off), p(b.read_bases), p(b.read_quals),
                           p(b.ins_gop), p(b.del_gop), p(b.gcp), p(b.hap_off), p(b.hap_bases), int(use_double),
                           int(fault), p(out), ec, em, leaks)
    return rc, out[:b.n_reads * b.n_haps], ec.value.decode(), em.value.decode(), (leaks[0], leaks[1])
