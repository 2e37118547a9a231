"""One config-5-shaped PDHMM call for ncu captures (GPU box):
    ncu --set full --clock-control none --import-source on -k regex:k_pdhmm3 -c 1 -o gpurun_out/pd python bench/profile_c5.py [reads]
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gkl_b200 import synth
from gkl_b200.pdhmm import IntelPDHMM
from gkl_b200.pdhmm_batch import PdhmmBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
reads, haps = synth.config5(n, 128)
ops = PdhmmBatch.operands(reads, haps)
hmm = IntelPDHMM()
assert hmm.load()
hmm.initialize(None)
out = np.zeros(len(reads) * len(haps))
for _ in range(3):
    hmm.compute_cross(ops, out)
print(hmm.kernel_name(), hmm.time_runs(2), "ms per run")
hmm.done()
