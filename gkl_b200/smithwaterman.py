"""Host-side mirror of GKL's Smith-Waterman operator interface (reference
src/main/java/com/intel/gkl/smithwaterman/IntelSmithWaterman.java and the gatk-native-bindings types it implements),
over the C-ABI of include/gklb_sw.h.

    sw = IntelSmithWaterman(); sw.load()                                   # :77-111 -> bool
    res = sw.align(ref, alt, SWParameters(200, -150, -260, -11), SWOverhangStrategy.SOFTCLIP)   # :122-151
    res.cigar, res.alignment_offset
    sw.close()

``align_batch`` is the batched form the device needs (the reference aligns one pair per JNI call): same validation
per pair, one kernel launch for all of them.
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import native
from .pairhmm import IllegalArgumentException, NullPointerException, OutOfMemoryError

MAX_SW_SEQUENCE_LENGTH = 32 * 1024 - 1   # IntelSmithWaterman.java:53
MAXIMUM_SW_MATCH_VALUE = 64 * 1024       # :55

SW_EXPORTS = ["gklb_sw_init", "gklb_sw_align_batch", "gklb_sw_align", "gklb_sw_done", "gklb_sw_last_stats",
              "gklb_sw_time_runs"]


class SWOverhangStrategy(enum.Enum):  # byte codes of IntelSmithWaterman.getStrategy (:153-170)
    SOFTCLIP = 9
    INDEL = 10
    LEADING_INDEL = 11
    IGNORE = 12


@dataclass
class SWParameters:  # gatk-native-bindings SWParameters
    matchValue: int
    mismatchPenalty: int
    gapOpenPenalty: int
    gapExtendPenalty: int


@dataclass
class SWNativeAlignerResult:
    cigar: str
    alignment_offset: int


class _SwBatch(C.Structure):
    _fields_ = [("n", C.c_int32), ("seq1", C.c_void_p), ("seq1_off", C.c_void_p), ("seq2", C.c_void_p),
                ("seq2_off", C.c_void_p), ("match", C.c_int32), ("mismatch", C.c_int32), ("open", C.c_int32),
                ("extend", C.c_int32), ("strategy", C.c_int32)]


class SwStats(C.Structure):
    _fields_ = [("pairs", C.c_int64), ("cells", C.c_int64), ("h2d_ms", C.c_float), ("kernel_ms", C.c_float),
                ("d2h_ms", C.c_float), ("kernel_launches", C.c_int32), ("warps", C.c_int32)]


def _lib():
    l = native.lib()
    l.gklb_sw_align_batch.argtypes = [C.POINTER(_SwBatch), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    l.gklb_sw_align.argtypes = [C.c_int32] * 4 + [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                                  C.c_int32, C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]
    l.gklb_sw_last_stats.argtypes = [C.POINTER(SwStats)]
    l.gklb_sw_time_runs.argtypes = [C.c_int, C.POINTER(C.c_float)]
    return l


def _raise(rc: int):
    msg = native.lib().gklb_last_error().decode(errors="replace")
    if rc == native.ERR_OOM:
        raise OutOfMemoryError("Memory allocation failed")          # IntelSmithWaterman.java:146-148
    if rc == native.ERR_INVALID:
        raise IllegalArgumentException("Ran into invalid argument issue: " + msg)  # :149-151
    raise native.GklbError(rc, msg)


def pack(seqs: Sequence[bytes]):
    """Concatenate byte strings into (uint8 arena, int64 offsets[n + 1])."""
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    arena = np.frombuffer(b"".join(bytes(s) for s in seqs), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, np.uint8)
    return arena, off


class IntelSmithWaterman:
    def __init__(self):
        self._loaded = False

    def load(self, tmp_dir=None) -> bool:
        """True when the library loads and an sm_100 device is usable (there is no CPU path); also runs initNative."""
        try:
            if native.device_count() <= 0:
                return False
            rc = _lib().gklb_sw_init()
        except (OSError, FileNotFoundError):
            return False
        self._loaded = rc == 0
        return self._loaded

    @staticmethod
    def _check(ref, alt, parameters, strategy):  # IntelSmithWaterman.java:122-145
        if ref is None:
            raise NullPointerException("Reference data array is null.")
        if alt is None:
            raise NullPointerException("Alternate data array is null.")
        if parameters is None:
            raise NullPointerException("Parameter structure is null.")
        if strategy is None:
            raise NullPointerException("OverhangStrategy is null.")
        if len(ref) <= 0 or len(alt) <= 0:
            raise IllegalArgumentException("Cannot align empty sequences")
        if len(ref) > MAX_SW_SEQUENCE_LENGTH or len(alt) > MAX_SW_SEQUENCE_LENGTH:
            raise IllegalArgumentException(f"Sequences exceed maximum length of {MAX_SW_SEQUENCE_LENGTH} bytes")
        if parameters.matchValue > MAXIMUM_SW_MATCH_VALUE:
            raise IllegalArgumentException(f"Match value parameter exceed maximum value of {MAXIMUM_SW_MATCH_VALUE}")
        code = strategy.value if isinstance(strategy, SWOverhangStrategy) else int(strategy)
        if code < 9 or code > 12:
            raise IllegalArgumentException("Strategy is invalid.")
        return code

    def align(self, refArray, altArray, parameters: SWParameters, overhangStrategy) -> SWNativeAlignerResult:
        code = self._check(refArray, altArray, parameters, overhangStrategy)
        ref = np.frombuffer(bytes(refArray), dtype=np.uint8)
        alt = np.frombuffer(bytes(altArray), dtype=np.uint8)
        cigar = np.zeros(2 * max(len(ref), len(alt)), dtype=np.uint8)  # :135
        count, offset = C.c_uint32(0), C.c_int32(0)
        rc = _lib().gklb_sw_align(parameters.matchValue, parameters.mismatchPenalty, parameters.gapOpenPenalty,
                                  parameters.gapExtendPenalty, ref.ctypes.data, alt.ctypes.data, len(ref), len(alt),
                                  code, cigar.ctypes.data, len(cigar), C.byref(count), C.byref(offset))
        if rc:
            _raise(rc)
        text = bytes(cigar).decode("utf-8").replace("\x00", " ").strip()  # new String(cigar, UTF_8).trim() (:153)
        return SWNativeAlignerResult(text, offset.value)

    def align_batch(self, refs: Sequence[bytes], alts: Sequence[bytes], parameters: SWParameters, overhangStrategy):
        """Returns (list of CIGAR strings, int32 offsets)."""
        if refs is None or alts is None:
            raise NullPointerException("Reference data array is null.")
        if len(refs) != len(alts):
            raise IllegalArgumentException("refs and alts differ in length")
        code = None
        for r, a in zip(refs, alts):
            code = self._check(r, a, parameters, overhangStrategy)
        if not refs:
            return [], np.zeros(0, dtype=np.int32)
        s1, o1 = pack(refs)
        s2, o2 = pack(alts)
        return self.align_packed(s1, o1, s2, o2, parameters, code)

    def align_packed_raw(self, s1, o1, s2, o2, parameters: SWParameters, code: int):
        """The C-ABI call alone: returns (cigar rows uint8[n, pitch], cigar lengths int32[n], offsets int32[n])."""
        n = len(o1) - 1
        pitch = int(2 * max(int(np.diff(o1).max()), int(np.diff(o2).max())))
        cig = np.zeros((n, pitch), dtype=np.uint8)
        clen = np.zeros(n, dtype=np.int32)
        offs = np.zeros(n, dtype=np.int32)
        b = _SwBatch(n, s1.ctypes.data, o1.ctypes.data, s2.ctypes.data, o2.ctypes.data, parameters.matchValue,
                     parameters.mismatchPenalty, parameters.gapOpenPenalty, parameters.gapExtendPenalty, int(code))
        rc = _lib().gklb_sw_align_batch(C.byref(b), cig.ctypes.data, pitch, clen.ctypes.data, offs.ctypes.data)
        if rc:
            _raise(rc)
        return cig, clen, offs

    def align_packed(self, s1, o1, s2, o2, parameters: SWParameters, code: int):
        cig, clen, offs = self.align_packed_raw(s1, o1, s2, o2, parameters, code)
        return [bytes(cig[k, :clen[k]]).decode("ascii") for k in range(len(clen))], offs

    def stats(self) -> SwStats:
        st = SwStats()
        _lib().gklb_sw_last_stats(C.byref(st))
        return st

    def time_runs(self, iters: int) -> float:
        ms = C.c_float(0)
        rc = _lib().gklb_sw_time_runs(iters, C.byref(ms))
        if rc:
            _raise(rc)
        return ms.value

    def close(self) -> None:  # IntelSmithWaterman.close -> doneNative
        _lib().gklb_sw_done()
        self._loaded = False
