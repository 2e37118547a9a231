"""Config 4 (1 M reads of length 150 x 256 haplotypes) through the process-global C-ABI surface -- the one the JNI
layer calls -- with GKLB_DEVICES=all: one process, one engine and one host thread per GPU, reads sharded.

    GKLB_DEVICES=all python bench/config4_inprocess.py [--reads 1000000] [--haps 256] [--out gpurun_out/config4.json]
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import oracle  # noqa: E402  (checker only)
from gkl_b200 import native, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--haps", type=int, default=256)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--out", default="gpurun_out/config4.json")
    a = ap.parse_args()
    os.environ.setdefault("GKLB_DEVICES", "all")
    t0 = time.time()
    b = synth.config4(a.reads, a.haps)
    gen_s = time.time() - t0
    n_dev = native.global_init(False, 1)
    out = np.empty(b.n_reads * b.n_haps, dtype=np.float64)
    native.global_compute(b, out)  # warm-up: allocations on every device
    best = 1e9
    for _ in range(a.reps):
        t0 = time.perf_counter()
        native.global_compute(b, out)
        best = min(best, time.perf_counter() - t0)
    st = native.global_stats()
    # parity on a stratified read sample (every 1 % of the reads) against GKL's own code / the restatement
    idx = np.linspace(0, b.n_reads - 1, 100).astype(int)
    fn = oracle.ref_pairhmm if oracle.ref_available() else oracle.port_pairhmm
    err = 0.0
    for r in idx:
        ref = fn(b.read_slice(int(r), int(r) + 1), threads=1)[0]
        got = out[int(r) * b.n_haps:(int(r) + 1) * b.n_haps]
        err = max(err, float(np.max(np.abs(got - ref) / np.abs(ref))))
    res = {"workload": f"config4: {b.n_reads} reads (len 150) x {b.n_haps} haplotypes, one process, GKLB_DEVICES",
           "devices": n_dev, "cells": b.cells(), "pairs": int(st.pairs), "fallback_pairs": int(st.fallback_pairs),
           "e2e_s": best, "e2e_gcups": b.cells() / best / 1e9,
           "phases_ms_max_over_devices": {"h2d_pack": st.h2d_ms, "kernels": st.kernel_ms, "d2h": st.d2h_ms},
           "kernels_gcups": b.cells() / st.kernel_ms / 1e6, "max_rel_err_100_read_sample": err,
           "host_generation_s": gen_s}
    native.global_done()
    print(json.dumps(res))
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(res) + "\n")


if __name__ == "__main__":
    main()
