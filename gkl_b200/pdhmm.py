"""Host-side mirror of GKL's PDHMM operator interface (reference
src/main/java/com/intel/gkl/pdhmm/IntelPDHMM.java), over the C-ABI of include/gklb_pdhmm.h.

    hmm = IntelPDHMM(); hmm.load(); hmm.initialize(PDHMMNativeArguments())
    out = hmm.computePDHMM(hap_bases, hap_pdbases, read_bases, read_qual, read_ins_qual, read_del_qual, gcp,
                           hap_lengths, read_lengths, batchSize, maxHapLength, maxReadLength)     # :163-204
    hmm.computeLikelihoods(readDataArray, haplotypeDataArray, likelihoodArray)                     # :92-121
    hmm.done()
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import native
from .pairhmm import IllegalArgumentException, NullPointerException, OutOfMemoryError
from .pdhmm_batch import PdhmmBatch


class _PdBatch(C.Structure):
    _fields_ = [("n", C.c_int64), ("max_hap", C.c_int32), ("max_read", C.c_int32)] + \
               [(k, C.c_void_p) for k in ("hap_bases", "hap_pdbases", "read_bases", "read_qual", "read_ins_qual",
                                          "read_del_qual", "gcp", "hap_lengths", "read_lengths")]


class PdStats(C.Structure):
    _fields_ = [("pairs", C.c_int64), ("cells", C.c_int64), ("h2d_ms", C.c_float), ("kernel_ms", C.c_float),
                ("d2h_ms", C.c_float), ("kernel_launches", C.c_int32)]


PD_EXPORTS = ["gklb_pdhmm_init", "gklb_pdhmm_compute", "gklb_pdhmm_compute_cross", "gklb_pdhmm_done",
              "gklb_pdhmm_last_stats", "gklb_pdhmm_time_runs", "gklb_pdhmm_table", "gklb_pdhmm_kernel_name"]


@dataclass
class PDReadDataHolder:  # gatk-native-bindings ReadDataHolder (pdhmm/JavaData.h:165-174)
    readBases: bytes
    readQuals: bytes
    insertionGOP: bytes
    deletionGOP: bytes
    overallGCP: bytes


@dataclass
class PDHaplotypeDataHolder:
    haplotypeBases: bytes
    haplotypePDBases: bytes


@dataclass
class PDHMMNativeArguments:  # IntelPDHMM.java:79-89; ordinals of pdhmm-implementation.h:45-58
    maxNumberOfThreads: int = 1
    avxLevel: int = 0        # FASTEST_AVAILABLE
    openMPSetting: int = 0   # FASTEST_AVAILABLE
    maxMemoryInMB: int = 512


def _lib():
    l = native.lib()
    l.gklb_pdhmm_init.argtypes = [C.c_int] * 4
    l.gklb_pdhmm_compute.argtypes = [C.POINTER(_PdBatch), C.c_void_p]
    l.gklb_pdhmm_compute_cross.argtypes = [C.POINTER(_PdBatch), C.c_int32, C.c_int32, C.c_void_p]
    l.gklb_pdhmm_last_stats.argtypes = [C.POINTER(PdStats)]
    l.gklb_pdhmm_time_runs.argtypes = [C.c_int, C.POINTER(C.c_float)]
    l.gklb_pdhmm_table.restype = C.c_void_p
    l.gklb_pdhmm_kernel_name.restype = C.c_char_p
    l.gklb_pdhmm_kernel_name.argtypes = []
    l.gklb_pdhmm_table.argtypes = [C.c_int, C.POINTER(C.c_int)]
    return l


def tables() -> dict:
    out = {}
    for which, name in enumerate(("q2err", "mm")):
        n = C.c_int(0)
        p = _lib().gklb_pdhmm_table(which, C.byref(n))
        out[name] = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n.value,)).copy()
    return out


def _struct(b: PdhmmBatch, n: int) -> _PdBatch:
    s = _PdBatch()
    s.n, s.max_hap, s.max_read = n, b.max_hap, b.max_read
    for k in ("hap_bases", "hap_pdbases", "read_bases", "read_qual", "read_ins_qual", "read_del_qual", "gcp",
              "hap_lengths", "read_lengths"):
        setattr(s, k, getattr(b, k).ctypes.data)
    return s


def _raise(rc: int):
    msg = native.lib().gklb_last_error().decode(errors="replace")
    if rc == native.ERR_OOM:
        raise OutOfMemoryError("OutOfMemory exception thrown from native pdhmm function call " + msg)
    if rc == native.ERR_INVALID:
        raise IllegalArgumentException("IllegalArgument exception thrown from native pdhmm function call " + msg)
    raise native.GklbError(rc, msg)


class IntelPDHMM:
    def __init__(self):
        self._init = False

    def load(self, tmp_dir=None) -> bool:
        try:
            return native.device_count() > 0
        except (OSError, FileNotFoundError):
            return False

    def initialize(self, args: Optional[PDHMMNativeArguments] = None) -> None:
        if args is None:
            args = PDHMMNativeArguments()
        rc = _lib().gklb_pdhmm_init(args.openMPSetting, args.maxNumberOfThreads, args.avxLevel, args.maxMemoryInMB)
        if rc:
            _raise(rc)
        self._init = True

    def computePDHMM(self, hap_bases, hap_pdbases, read_bases, read_qual, read_ins_qual, read_del_qual, gcp,
                     hap_lengths, read_lengths, batchSize: int, maxHapLength: int, maxReadLength: int) -> np.ndarray:
        arrays = dict(hap_bases=hap_bases, hap_pdbases=hap_pdbases, read_bases=read_bases, read_qual=read_qual,
                      read_ins_qual=read_ins_qual, read_del_qual=read_del_qual, gcp=gcp, hap_lengths=hap_lengths,
                      read_lengths=read_lengths)
        expect = {k: maxHapLength * batchSize for k in ("hap_bases", "hap_pdbases")}
        expect.update({k: maxReadLength * batchSize for k in ("read_bases", "read_qual", "read_ins_qual", "read_del_qual", "gcp")})
        expect.update(hap_lengths=batchSize, read_lengths=batchSize)
        for k, v in arrays.items():  # IntelPDHMM.checkArraySize (:139-150)
            if v is None:
                raise NullPointerException(f"{k} must not be null.")
            if len(v) != expect[k]:
                raise IllegalArgumentException(f"Array {k} has size {len(v)}, but expected size is {expect[k]}.")
        if batchSize <= 0:
            raise IllegalArgumentException("batchSize must be greater than 0.")
        if maxHapLength <= 0 or maxReadLength <= 0:
            raise IllegalArgumentException("maxHapLength / maxReadLength must be greater than 0. Cannot perform PDHMM on empty sequence")
        i8 = lambda a: np.ascontiguousarray(a, dtype=np.int8)
        b = PdhmmBatch(i8(hap_bases), i8(hap_pdbases), i8(read_bases), i8(read_qual), i8(read_ins_qual), i8(read_del_qual),
                       i8(gcp), np.ascontiguousarray(hap_lengths, dtype=np.int64),
                       np.ascontiguousarray(read_lengths, dtype=np.int64), maxHapLength, maxReadLength)
        return self.compute_batch(b)

    def compute_batch(self, b: PdhmmBatch) -> np.ndarray:
        out = np.empty(b.n, dtype=np.float64)
        s = _struct(b, b.n)
        rc = _lib().gklb_pdhmm_compute(C.byref(s), out.ctypes.data)
        if rc:
            _raise(rc)
        return out

    def computeLikelihoods(self, readDataArray: Sequence[PDReadDataHolder], haplotypeDataArray: Sequence[PDHaplotypeDataHolder],
                           likelihoodArray: np.ndarray) -> None:
        if readDataArray is None or haplotypeDataArray is None or likelihoodArray is None:
            raise NullPointerException("One or more input arrays are null.")
        if len(likelihoodArray) != len(readDataArray) * len(haplotypeDataArray):
            raise IllegalArgumentException("likelihoodArray length must be equal to readDataArray length * haplotypeDataArray length")
        # operands once (haplotype h at h * max_hap, read r at r * max_read); the cross product is device-side
        b = PdhmmBatch.operands(
            [(r.readBases, r.readQuals, r.insertionGOP, r.deletionGOP, r.overallGCP) for r in readDataArray],
            [(h.haplotypeBases, h.haplotypePDBases) for h in haplotypeDataArray])
        self.compute_cross(b, likelihoodArray)

    def compute_cross(self, b: PdhmmBatch, out: np.ndarray) -> None:
        """gklb_pdhmm_compute_cross on prepared operands (PdhmmBatch.operands): out[r * H + h]."""
        s = _struct(b, 0)
        rc = _lib().gklb_pdhmm_compute_cross(C.byref(s), len(b.read_lengths), len(b.hap_lengths), out.ctypes.data)
        if rc:
            _raise(rc)

    def stats(self) -> PdStats:
        st = PdStats()
        _lib().gklb_pdhmm_last_stats(C.byref(st))
        return st

    def time_runs(self, iters: int) -> float:
        ms = C.c_float(0)
        rc = _lib().gklb_pdhmm_time_runs(iters, C.byref(ms))
        if rc:
            _raise(rc)
        return ms.value

    def kernel_name(self) -> str:
        """The kernel that carried the last compute call (gklb_pdhmm_kernel_name)."""
        return _lib().gklb_pdhmm_kernel_name().decode()

    def done(self) -> None:
        _lib().gklb_pdhmm_done()
        self._init = False
