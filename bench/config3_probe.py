"""Config 3 probe (GPU box): kernel time per region, and of jobs of K regions, with the plan of each."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gkl_b200 import native, synth

regions = synth.config3(32, seed=3)
cells = sum(r.cells() for r in regions)
e = native.Engine(0, False)
tot = 0.0
for i, r in enumerate(regions):
    e.stage(r); e.run(); e.synchronize()
    ms = e.time_runs(5)
    tot += ms
    st = e.stats()
    if i < 3:
        print(f"region {i}: {r.n_reads}x{r.n_haps} cells={r.cells():.3e} ms={ms:.3f} sweep_ms={st.sweep_ms:.3f} classes={st.n_classes}")
        print(e.plan_info())
print(f"per-region kernels: {tot:.2f} ms = {cells / tot / 1e6:.0f} GCUPS")
for K in (1, 2, 4, 8, 16, 32):
    t = 0.0
    sw = 0.0
    for i in range(0, 32, K):
        e.stage_multi(regions[i:i + K]); e.run(); e.synchronize()
        t += e.time_runs(3)
        sw += e.stats().sweep_ms
    print(f"jobs of {K:2d} regions: kernels {t:.2f} ms = {cells / t / 1e6:.0f} GCUPS, sweep launches {sw:.2f} ms = {cells / sw / 1e6:.0f} GCUPS")
    if K in (8, 32):
        print(e.plan_info())
t0 = time.perf_counter()
outs = e.compute_multi(regions)
t1 = time.perf_counter()
outs = e.compute_multi(regions, outs)
t2 = time.perf_counter()
st = e.stats()
print(f"compute_multi(32): {1e3 * (t2 - t1):.2f} ms (first {1e3 * (t1 - t0):.2f}) phases h2d {st.h2d_ms:.2f} kernels {st.kernel_ms:.2f} d2h {st.d2h_ms:.2f}")
