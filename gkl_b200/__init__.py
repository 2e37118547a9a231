"""B200-native PairHMM engine behind GKL's PairHMM native-binding surface.

The package holds only what the hot path needs: the CUDA kernels and C-ABI under ``csrc/``
(built into ``lib/libgkl_pairhmm.so``), the ctypes loader (``native``), and the host-side mirror of
GKL's ``IntelPairHmm`` operator interface (``pairhmm``).
"""
from .batch import PairHmmBatch  # noqa: F401
from .pairhmm import (HaplotypeDataHolder, IntelPairHmm, PairHMMNativeArguments,  # noqa: F401
                      ReadDataHolder)

__all__ = ["PairHmmBatch", "IntelPairHmm", "ReadDataHolder", "HaplotypeDataHolder", "PairHMMNativeArguments"]
