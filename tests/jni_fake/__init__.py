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
        subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-pthread", "-o", str(SO), str(src), "-ldl"], check=True)


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


def pairhmm_threads(lib_path, batches, n_threads: int, rounds: int = 1):
    """n_threads concurrent callers, each with its own JNIEnv and IntelPairHmm instance, share `batches` round-robin.
    Returns (failed_calls, [likelihood arrays], last exception message, (leaked local refs, leaked pins))."""
    l = _lib()
    n = len(batches)
    outs = [np.zeros(max(1, b.n_reads * b.n_haps), dtype=np.float64) for b in batches]
    arr_i = lambda vals: (C.c_int * n)(*vals)
    arr_p = lambda arrays: (C.c_void_p * n)(*[a.ctypes.data for a in arrays])
    em = C.create_string_buffer(256)
    leaks = (C.c_long * 2)()
    rc = l.fakejvm_pairhmm_mt(str(lib_path).encode(), int(n_threads), n, arr_i([b.n_reads for b in batches]),
                              arr_i([b.n_haps for b in batches]), arr_p([b.read_off for b in batches]),
                              arr_p([b.read_bases for b in batches]), arr_p([b.read_quals for b in batches]),
                              arr_p([b.ins_gop for b in batches]), arr_p([b.del_gop for b in batches]),
                              arr_p([b.gcp for b in batches]), arr_p([b.hap_off for b in batches]),
                              arr_p([b.hap_bases for b in batches]), arr_p(outs), int(rounds), leaks, em)
    return rc, [o[:b.n_reads * b.n_haps] for o, b in zip(outs, batches)], em.value.decode(), (leaks[0], leaks[1])


def pdhmm(lib_path, b, object_api=False, n_reads=0, n_haps=0, fault=0):
    """b: gkl_b200.pdhmm_batch.PdhmmBatch -- the flat batch (object_api False) or strided operands
    (object_api True: n_reads reads, n_haps haplotypes).  Returns (rc, out, exception class, message, leaks)."""
    l = _lib()
    n = n_reads * n_haps if object_api else b.n
    out = np.zeros(max(1, n), dtype=np.float64)
    ec, em = C.create_string_buffer(256), C.create_string_buffer(256)
    leaks = (C.c_long * 2)()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = l.fakejvm_pdhmm(str(lib_path).encode(), int(object_api), C.c_longlong(b.n), n_reads, n_haps, int(b.max_hap),
                         int(b.max_read), p(b.hap_bases), p(b.hap_pdbases), p(b.read_bases), p(b.read_qual),
                         p(b.read_ins_qual), p(b.read_del_qual), p(b.gcp), p(b.hap_lengths), p(b.read_lengths), int(fault),
                         p(out), ec, em, leaks)
    return rc, out[:n], ec.value.decode(), em.value.decode(), (leaks[0], leaks[1])


def smithwaterman(lib_path, ref: bytes, alt: bytes, params, strategy: int, fault=0):
    """One IntelSmithWaterman.alignNative call.  Returns (rc, cigar bytes (2 * max(len) like the Java wrapper
    allocates), offset, exception class, message, leaks)."""
    l = _lib()
    cap = 2 * max(len(ref), len(alt), 1)
    cigar = C.create_string_buffer(cap)
    off = C.c_int(0)
    ec, em = C.create_string_buffer(256), C.create_string_buffer(256)
    leaks = (C.c_long * 2)()
    rc = l.fakejvm_sw(str(lib_path).encode(), ref, len(ref), alt, len(alt), *[int(v) for v in params], int(strategy),
                      int(fault), cigar, cap, C.byref(off), ec, em, leaks)
    return rc, cigar.raw, off.value, ec.value.decode(), em.value.decode(), (leaks[0], leaks[1])
