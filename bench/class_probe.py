"""Per-class efficiency probe (GPU box): sweep GCUPS of uniform-length batches, through the single-class kernel and
through the multi-class kernel (forced by adding a handful of reads of another length)."""
import os, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gkl_b200 import native, synth

e = native.Engine(0, False)
for L in (36, 48, 64, 72, 88, 101, 104, 128, 150, 160, 200, 250):
    n = max(512, int(1.0e10 / (L * 300 * 128)) // 64 * 64)
    b = synth.config2(n, 128, L)
    e.stage(b); e.run(); e.synchronize()
    ms = e.time_runs(3)
    st = e.stats()
    row = f"L={L:3d} reads={n:6d} {e.sweep_kernel():24s} sweep {b.cells() / st.sweep_ms / 1e6:6.0f} GCUPS  all kernels {b.cells() / ms / 1e6:6.0f}"
    # the same reads plus 64 reads of another class -> multi-class launch
    rng = np.random.default_rng(1)
    haps = [b.hap_bases[b.hap_off[h]:b.hap_off[h + 1]] for h in range(b.n_haps)]
    lens = np.concatenate([np.full(n, L), np.full(64, 40 if L > 60 else 100)]).astype(np.int64)
    m = synth._assemble(haps, synth._reads_from_panel(rng, haps, lens, empirical=True))
    e.stage(m); e.run(); e.synchronize()
    st = e.stats()
    row += f"   | mega: sweep {m.cells() / st.sweep_ms / 1e6:6.0f} GCUPS ({e.sweep_kernel()})"
    print(row, flush=True)
