"""TEST INFRASTRUCTURE ONLY.

ctypes access to the two CPU checkers:

* ``port_*``  -- oracle/libgklb_oracle.so, the CPU restatement written in this repo
  (pairhmm_oracle.c / pdhmm_oracle.c; ``make -C oracle port``);
* ``ref_*``   -- oracle/_ref/libgkl_ref.so, GKL's own unmodified translation units compiled from
  /root/reference (``make -C oracle ref``; only buildable in the development container, the
  built .so travels to the GPU box).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` leg
may import this package.  The product (gkl_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
PORT_SO = HERE / "libgklb_oracle.so"
REF_SO = HERE / "_ref" / "libgkl_ref.so"
REFERENCE_ROOT = Path("/root/reference")

_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(ref: bool | None = None) -> None:
    """Compile the checkers (idempotent; make decides what is stale)."""
    targets = ["port"]
    if ref is None:
        ref = REFERENCE_ROOT.is_dir()
    if ref:
        targets.append("ref")
    subprocess.run(["make", "-s", "-C", str(HERE), *targets], check=True)


_port = None
_ref = None


def _load_port():
    global _port
    if _port is None:
        if not PORT_SO.exists():
            build(ref=False)
        lib = C.CDLL(str(PORT_SO))
        lib.gklport_pairhmm.restype = C.c_int
        lib.gklport_pairhmm.argtypes = [C.c_int, C.c_int, _i64p, _u8p, _u8p, _u8p, _u8p, _u8p, _i64p, _u8p,
                                        C.c_int, C.c_int, _f64p, C.c_void_p, C.POINTER(C.c_double)]
        lib.gklport_max_threads.restype = C.c_int
        for name, ty in (("gklport_ph2pr_f", C.c_float), ("gklport_mm_f", C.c_float),
                         ("gklport_ph2pr_d", C.c_double), ("gklport_mm_d", C.c_double)):
            getattr(lib, name).restype = C.POINTER(ty)
        _port = lib
    return _port


def ref_available() -> bool:
    return REF_SO.exists()


def _load_ref():
    global _ref
    if _ref is None:
        if not REF_SO.exists():
            if REFERENCE_ROOT.is_dir():
                build(ref=True)
            else:
                raise FileNotFoundError(f"{REF_SO} missing and /root/reference absent: run `make -C oracle ref` "
                                        "in the development container")
        lib = C.CDLL(str(REF_SO))
        lib.gklref_pairhmm.restype = C.c_int
        lib.gklref_pairhmm.argtypes = [C.c_int, C.c_int, _i64p, _u8p, _u8p, _u8p, _u8p, _u8p, _i64p, _u8p,
                                       C.c_int, C.c_int, C.c_int, _f64p, C.POINTER(C.c_int),
                                       C.POINTER(C.c_double)]
        lib.gklref_max_threads.restype = C.c_int
        lib.gklref_avx512_supported.restype = C.c_int
        _ref = lib
    return _ref


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def port_pairhmm(batch, use_double: bool = False, threads: int = 1):
    """CPU restatement.  Returns (log10 likelihoods[R*H], fallback flags[R*H], pair-loop seconds)."""
    lib = _load_port()
    n = batch.n_reads * batch.n_haps
    out = np.empty(n, dtype=np.float64)
    fb = np.zeros(n, dtype=np.uint8)
    secs = C.c_double(0)
    rc = lib.gklport_pairhmm(batch.n_reads, batch.n_haps, batch.read_off, batch.read_bases, batch.read_quals,
                             batch.ins_gop, batch.del_gop, batch.gcp, batch.hap_off, batch.hap_bases,
                             int(use_double), int(threads), out, fb.ctypes.data_as(C.c_void_p), C.byref(secs))
    if rc != 0:
        raise RuntimeError(f"gklport_pairhmm failed: {rc}")
    return out, fb, secs.value


def ref_pairhmm(batch, use_double: bool = False, threads: int = 1, engine: int = 0):
    """GKL's own compiled AVX / AVX-512 PairHMM driven by the loop of IntelPairHmm.cc:150-169.

    engine: 0 = GKL's CPUID dispatch, 1 = force AVX, 2 = force AVX-512.
    Returns (log10 likelihoods[R*H], used_avx512, pair-loop seconds)."""
    lib = _load_ref()
    n = batch.n_reads * batch.n_haps
    out = np.empty(n, dtype=np.float64)
    used = C.c_int(0)
    secs = C.c_double(0)
    rc = lib.gklref_pairhmm(batch.n_reads, batch.n_haps, batch.read_off, batch.read_bases, batch.read_quals,
                            batch.ins_gop, batch.del_gop, batch.gcp, batch.hap_off, batch.hap_bases,
                            int(use_double), int(threads), int(engine), out, C.byref(used), C.byref(secs))
    if rc != 0:
        raise RuntimeError(f"gklref_pairhmm failed: {rc}")
    return out, bool(used.value), secs.value


def ref_pairhmm_pairs(batch, pair_r: np.ndarray, pair_h: np.ndarray, threads: int = 1):
    """GKL's own compiled PairHMM (fp32, fp64 rerun under 1e-28, as IntelPairHmm.cc:150-169) on an explicit list of
    (read, haplotype) pairs of `batch`.  Returns (log10 likelihoods[len(pair_r)], seconds)."""
    lib = _load_ref()
    _i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    lib.gklref_pairhmm_pairs.restype = C.c_int
    lib.gklref_pairhmm_pairs.argtypes = [C.c_long, _i32, _i32, _i64p, _u8p, _u8p, _u8p, _u8p, _u8p, _i64p, _u8p, C.c_int,
                                         _f64p, C.POINTER(C.c_double)]
    pr = np.ascontiguousarray(pair_r, dtype=np.int32)
    ph = np.ascontiguousarray(pair_h, dtype=np.int32)
    out = np.empty(len(pr), dtype=np.float64)
    secs = C.c_double(0)
    rc = lib.gklref_pairhmm_pairs(len(pr), pr, ph, batch.read_off, batch.read_bases, batch.read_quals, batch.ins_gop,
                                  batch.del_gop, batch.gcp, batch.hap_off, batch.hap_bases, int(threads), out,
                                  C.byref(secs))
    if rc != 0:
        raise RuntimeError(f"gklref_pairhmm_pairs failed: {rc}")
    return out, secs.value


def ref_avx512_supported() -> bool:
    return bool(_load_ref().gklref_avx512_supported())


_i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")
_PD_ARGS = [_i8p] * 7 + [_f64p, C.c_int64, _i64p, _i64p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]


def _pd_call(fn, b, mode: int, threads: int):
    out = np.empty(b.n, dtype=np.float64)
    secs = C.c_double(0)
    rc = fn(b.hap_bases, b.hap_pdbases, b.read_bases, b.read_qual, b.read_ins_qual, b.read_del_qual, b.gcp, out, b.n,
            b.hap_lengths, b.read_lengths, int(b.max_read), int(b.max_hap), int(mode), int(threads), C.byref(secs))
    return out, rc, secs.value


def port_pdhmm(batch, carry_state: bool = True, threads: int = 1):
    """CPU restatement of the PDHMM (gkl_b200.pdhmm_batch.PdhmmBatch in).  carry_state=True is the serial
    reference path (and GATK's Java); False resets the column state per row like the AVX paths.
    Returns (log10 likelihoods[n], status, seconds)."""
    lib = _load_port()
    lib.gklport_pdhmm.restype = C.c_int
    lib.gklport_pdhmm.argtypes = _PD_ARGS
    return _pd_call(lib.gklport_pdhmm, batch, 1 if carry_state else 0, threads)


def ref_pdhmm(batch, level: int = 0, threads: int = 1):
    """GKL's own compiled PDHMM.  level: 0 fastest available, 1 scalar, 2 AVX2, 3 AVX-512.
    Returns (log10 likelihoods[n], reference status code, seconds)."""
    lib = _load_ref()
    lib.gklref_pdhmm.restype = C.c_int
    lib.gklref_pdhmm.argtypes = _PD_ARGS
    return _pd_call(lib.gklref_pdhmm, batch, level, threads)


def port_tables():
    """Host tables of the restatement (for bit-exact comparison with the engine's tables)."""
    lib = _load_port()
    mm_size = (255 * 256) // 2
    return {
        "ph2pr_f": np.ctypeslib.as_array(lib.gklport_ph2pr_f(), shape=(128,)).copy(),
        "mm_f": np.ctypeslib.as_array(lib.gklport_mm_f(), shape=(mm_size,)).copy(),
        "ph2pr_d": np.ctypeslib.as_array(lib.gklport_ph2pr_d(), shape=(128,)).copy(),
        "mm_d": np.ctypeslib.as_array(lib.gklport_mm_d(), shape=(mm_size,)).copy(),
    }


# ---- Smith-Waterman (smithwaterman/PairWiseSW.h) ----
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def _sw_call(fn, extra, seq1, off1, seq2, off2, params, strategy, threads):
    n = len(off1) - 1
    len1, len2 = np.diff(off1), np.diff(off2)
    pitch = int(2 * max(int(len1.max(initial=1)), int(len2.max(initial=1))))
    cig = np.zeros((n, pitch), dtype=np.uint8)
    clen = np.zeros(n, dtype=np.int32)
    offs = np.zeros(n, dtype=np.int32)
    secs = C.c_double(0)
    rc = fn(n, seq1, off1, seq2, off2, *[int(v) for v in params], int(strategy), *extra, int(threads),
            cig.ctypes.data_as(C.c_void_p), pitch, clen, offs, *([None] if extra else []), C.byref(secs))
    if rc != 0:
        raise MemoryError(f"smith-waterman oracle failed: {rc}")
    cigars = [bytes(cig[k, :clen[k]]).decode("ascii") for k in range(n)]
    return cigars, offs, secs.value


def port_sw(seq1, off1, seq2, off2, params, strategy: int, threads: int = 1):
    """CPU restatement (oracle/sw_oracle.c).  seq1/seq2: uint8 arenas, off1/off2: int64[n+1]; params = (match,
    mismatch, open, extend); strategy 9 SOFTCLIP, 10 INDEL, 11 LEADING_INDEL, 12 IGNORE.
    Returns (cigar strings, alignment offsets int32[n], seconds)."""
    lib = _load_port()
    lib.gklport_sw.restype = C.c_int
    lib.gklport_sw.argtypes = [C.c_int, _u8p, _i64p, _u8p, _i64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_void_p, C.c_int, _i32p, _i32p, C.POINTER(C.c_double)]
    return _sw_call(lib.gklport_sw, (), seq1, off1, seq2, off2, params, strategy, threads)


def ref_sw(seq1, off1, seq2, off2, params, strategy: int, threads: int = 1, engine: int = 0):
    """GKL's own compiled Smith-Waterman (engine: 0 its CPUID dispatch, 1 AVX2, 2 AVX-512), one
    runSWOnePairBT call per pair as IntelSmithWaterman.cc:72-121 makes them."""
    lib = _load_ref()
    lib.gklref_sw.restype = C.c_int
    lib.gklref_sw.argtypes = [C.c_int, _u8p, _i64p, _u8p, _i64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_int, C.c_void_p, C.c_int, _i32p, _i32p, C.c_void_p, C.POINTER(C.c_double)]
    return _sw_call(lib.gklref_sw, (int(engine),), seq1, off1, seq2, off2, params, strategy, threads)
