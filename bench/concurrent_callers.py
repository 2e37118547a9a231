"""Concurrent callers (GPU box): T host threads, each making the 32 synchronous config-3 calls through the global
surface (every call borrows an engine of its own from the pool).
    python bench/concurrent_callers.py
"""
import sys, time, threading
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gkl_b200 import native, synth
regions = synth.config3(32, seed=3)
cells = sum(r.cells() for r in regions)
native.global_init(False, 1)
outs = [native.global_compute(r) for r in regions]
for nt in (1, 2, 4, 8):
    def work():
        for r in regions: native.global_compute(r)
    best = 1e9
    for rep in range(3):
        th = [threading.Thread(target=work) for _ in range(nt)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        best = min(best, time.perf_counter() - t0)
    print(f"threads {nt}: {best*1e3:.1f} ms for {nt}x32 calls -> {nt*cells/best/1e9:.0f} GCUPS aggregate, {best*1e3/32:.3f} ms per call per thread", flush=True)
native.global_compute_multi(regions)
t0 = time.perf_counter(); native.global_compute_multi(regions); dt = time.perf_counter() - t0
print("one multi call of 32 regions: %.1f ms -> %.0f GCUPS" % (dt * 1e3, cells / dt / 1e9))
native.global_done()
