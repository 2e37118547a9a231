"""Measurement harness (GPU box): parity + GCUPS of every compiled kernel variant on a C2-shaped batch.

    python bench/sweep.py [--reads 4000] [--haps 128] [--iters 3] [--variants all|product|list:v1;v2]

Build the library with `make -C gkl_b200/csrc EXPERIMENTAL=1` to get the measurement-only variants.
Parity is checked against oracle/_ref (GKL's own AVX code) when present, else the oracle port.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import oracle  # noqa: E402  (measurement harness: the oracle is the checker only)
from gkl_b200 import native, synth  # noqa: E402

VARIANTS_H2 = ["f2,16,7,12,5", "h2,8,13,8,5", "h2,8,13,12,5", "h2,16,7,12,5", "h2,16,7,16,5", "h2,16,10,12,5",
               "h2,16,10,8,5", "h2,8,14,8,5", "h2,8,16,8,5", "h2,4,16,8,5"]
VARIANTS = [
    "f2,16,7,8,4", "f2,16,7,12,5", "f2,16,7,8,5", "f2,16,7,8,3", "f2,16,7,8,2", "f2,16,7,8,1", "f2,16,7,8,0", "f2,16,8,8,4", "f2,16,7,10,4", "f2,16,7,12,4",
    "f2,32,4,12,4", "f2,32,4,16,4", "f2,32,5,8,4",
    "f1,8,13,8,4", "f1,8,13,8,3", "f1,8,13,12,4", "f1,16,7,16,4",
    "d1,16,7,8,3", "d1,16,7,8,2",
]


def reference(b, threads):
    if oracle.ref_available():
        out, avx512, secs = oracle.ref_pairhmm(b, threads=threads)
        return out, secs, "reference(avx512)" if avx512 else "reference(avx)"
    out, _, secs = oracle.port_pairhmm(b, threads=threads)
    return out, secs, "port"


def rel_err(a, ref):
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), 1e-30)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=4000)
    ap.add_argument("--haps", type=int, default=128)
    ap.add_argument("--read-len", type=int, default=101)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--variants", default="all")
    ap.add_argument("--out", default="gpurun_out/sweep.jsonl")
    args = ap.parse_args()

    b = synth.config2(args.reads, args.haps, args.read_len)
    cells = b.cells()
    threads = oracle.host_threads()
    t0 = time.time()
    ref, cpu_secs, kind = reference(b, threads)
    print(f"# batch {b.n_reads}x{b.n_haps} cells={cells:.3e}  cpu {kind} {threads} threads: {cpu_secs:.3f}s "
          f"= {cells / cpu_secs / 1e9:.2f} GCUPS (wall {time.time() - t0:.1f}s)", flush=True)
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    rows = []
    if args.variants == "all":
        variants = [None] + VARIANTS
    elif args.variants == "h2":
        variants = [None] + VARIANTS_H2
    elif args.variants.startswith("list:"):   # e.g. list:f2,16,7,12,5;f2,16,7,12,6
        variants = [None] + [v for v in args.variants[5:].split(";") if v]
    else:
        variants = [None]
    for v in variants:
        if v is None:
            os.environ.pop("GKLB_FORCE_KERNEL", None)
        else:
            os.environ["GKLB_FORCE_KERNEL"] = v
        try:
            e = native.Engine(0, False)
            out = e.compute(b)
            st = e.stats()
            e.stage(b)
            e.run()
            e.synchronize()
            ms = e.time_runs(args.iters)
            err = rel_err(out, ref)
            row = {"variant": v or "default", "ms": ms, "gcups": cells / ms / 1e6, "max_rel_err": err,
                   "fallback": int(st.fallback_pairs), "launches": int(st.kernel_launches),
                   "e2e_ms": st.h2d_ms + st.kernel_ms + st.d2h_ms, "nan": int(np.isnan(out).sum())}
            e.close()
        except Exception as ex:  # keep sweeping
            row = {"variant": v or "default", "error": str(ex)}
        rows.append(row)
        print(json.dumps(row), flush=True)
        with open(args.out, "a") as f:
            f.write(json.dumps(row) + "\n")


if __name__ == "__main__":
    main()
