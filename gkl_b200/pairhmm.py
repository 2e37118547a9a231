"""Host-side mirror of GKL's PairHMM operator interface, over the C-ABI.

Same names, argument meaning and error behaviour as the Java shim it stands in for
(/root/reference/src/main/java/com/intel/gkl/pairhmm/IntelPairHmm.java and the
gatk-native-bindings types it implements):

    hmm = IntelPairHmm()
    hmm.load()                                     # IntelPairHmm.java:65-82  -> bool
    hmm.initialize(PairHMMNativeArguments(...))    # :85-119
    hmm.computeLikelihoods(reads, haps, out)       # :130-147   out[r * len(haps) + h] = log10 L
    hmm.done()                                     # :153-155

In a JVM the unchanged Java class calls the three ``Java_com_intel_gkl_pairhmm_IntelPairHmm_*``
symbols of libgkl_pairhmm.so (csrc/jni_pairhmm.cc), which marshal into exactly the flat batch
built here.  There is no JVM in this environment, so tests and benchmarks enter at this level.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import native
from .batch import PairHmmBatch


class NullPointerException(Exception):
    """IntelPairHmm.computeLikelihoods: 'Input is null' (IntelPairHmm.java:134-136)."""


class OutOfMemoryError(Exception):
    """'Memory allocation failed' (IntelPairHmm.java:140-142)."""


class IllegalArgumentException(Exception):
    """'Ran into invalid argument issue' (IntelPairHmm.java:143-145)."""


@dataclass
class ReadDataHolder:  # gatk-native-bindings: byte[] fields read by JavaData (pairhmm/JavaData.h:55-62)
    readBases: bytes
    readQuals: bytes
    insertionGOP: bytes
    deletionGOP: bytes
    overallGCP: bytes


@dataclass
class HaplotypeDataHolder:
    haplotypeBases: bytes


@dataclass
class PairHMMNativeArguments:
    useDoublePrecision: bool = False
    maxNumberOfThreads: int = 1


class IntelPairHmm:
    def __init__(self) -> None:
        self._engine: Optional[native.Engine] = None
        self._loaded = False
        self.device = 0

    def load(self, tmp_dir=None) -> bool:
        """True when the native library is present and a B200-class device is usable.  Returning
        False is the only 'fallback': the caller (GATK) then uses its own Java PairHMM."""
        try:
            self._loaded = native.device_count() > 0
        except (OSError, FileNotFoundError):
            self._loaded = False
        return self._loaded

    def initialize(self, args: Optional[PairHMMNativeArguments] = None) -> None:
        if args is None:  # IntelPairHmm.java:86-90
            args = PairHMMNativeArguments(False, 1)
        self.done()
        self._engine = native.Engine(self.device, bool(args.useDoublePrecision))

    def computeLikelihoods(self, readDataArray: Sequence[ReadDataHolder],
                           haplotypeDataArray: Sequence[HaplotypeDataHolder], likelihoodArray: np.ndarray) -> None:
        if readDataArray is None or haplotypeDataArray is None or likelihoodArray is None:
            raise NullPointerException("Input is null.")
        if self._engine is None:
            raise RuntimeError("initialize() has not been called")
        if len(readDataArray) == 0 or len(haplotypeDataArray) == 0:
            return  # the native loop simply runs zero iterations (IntelPairHmm.cc:150-169)
        for r in readDataArray:
            if r is None or None in (r.readBases, r.readQuals, r.insertionGOP, r.deletionGOP, r.overallGCP):
                raise NullPointerException("Input is null.")
        batch = PairHmmBatch.from_lists(
            [r.readBases for r in readDataArray], [r.readQuals for r in readDataArray],
            [r.insertionGOP for r in readDataArray], [r.deletionGOP for r in readDataArray],
            [r.overallGCP for r in readDataArray], [h.haplotypeBases for h in haplotypeDataArray])
        n = batch.n_reads * batch.n_haps
        if likelihoodArray.dtype != np.float64 or not likelihoodArray.flags.c_contiguous or likelihoodArray.size < n:
            raise IllegalArgumentException("Ran into invalid argument issue.")
        try:
            self._engine.compute(batch, likelihoodArray)
        except native.GklbError as e:
            if e.code == native.ERR_OOM:
                raise OutOfMemoryError("Memory allocation failed.") from e
            if e.code == native.ERR_INVALID:
                raise IllegalArgumentException("Ran into invalid argument issue.") from e
            raise

    def compute_batch(self, batch: PairHmmBatch, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Flat-batch entry (what the JNI layer hands to the C-ABI after marshalling)."""
        if self._engine is None:
            raise RuntimeError("initialize() has not been called")
        return self._engine.compute(batch, out)

    def done(self) -> None:
        if self._engine is not None:
            self._engine.close()
            self._engine = None
