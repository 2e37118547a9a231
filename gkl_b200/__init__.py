"""B200-native PairHMM engine behind GKL's PairHMM native-binding surface.

The package holds only what the hot path needs: the CUDA kernels and C-ABI under ``csrc/``
(built into ``libgkl_pairhmm.so``), the ctypes loader, and the host-side mirror of GKL's
``IntelPairHmm`` operator interface.
"""
from .batch import PairHmmBatch  # noqa: F401

__all__ = ["PairHmmBatch"]
