"""Throughput by read length (GPU box): synth.config2(n, 128, L) at ~3.8e10 cells per length, kernels incl. the rerun.
    python bench/len_sweep.py > profiles/r2_read_length_sweep.jsonl
"""
import sys, json, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gkl_b200 import native, synth
eng = native.Engine(0, False)
rows = []
for L in [36, 50, 76, 90, 101, 110, 125, 140, 151, 175, 200, 250, 300]:
    n = max(500, int(10000 * 101 / L))
    b = synth.config2(n, 128, L)
    eng.compute(b)
    best = 1e9
    for _ in range(3):
        eng.compute(b)
        st = eng.stats()
        best = min(best, st.kernel_ms)
    info = eng.plan_info() if hasattr(eng, 'plan_info') else None
    rows.append({"read_len": L, "reads": n, "cells": b.cells(), "kernel_ms": best, "gcups": b.cells() / best / 1e6,
                 "fallback_frac": st.fallback_pairs / (b.n_reads * b.n_haps), "kernel": eng.sweep_kernel()})
    print(json.dumps(rows[-1]), flush=True)
