"""PDHMM (config 5) measurement on the GPU box: GCUPS of the CUDA path (kernel-resident and end to end through
the operator interface), GKL's own AVX-512 / scalar PDHMM on the host, and parity on a pair sample.

    python bench/pdhmm_bench.py [--reads 10000] [--haps 128] [--iters 3] [--out gpurun_out/pdhmm_bench.json]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import oracle  # noqa: E402  (checker / CPU baseline only)
from gkl_b200 import pdhmm_batch as pb, synth  # noqa: E402
from gkl_b200.pdhmm import IntelPDHMM, PDHaplotypeDataHolder, PDReadDataHolder  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000)
    ap.add_argument("--haps", type=int, default=128)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--cpu-reads", type=int, default=150)
    ap.add_argument("--read-len", type=int, default=101)
    ap.add_argument("--out", default="gpurun_out/pdhmm_bench.json")
    a = ap.parse_args()
    reads, haps = synth.config5(a.reads, a.haps, a.read_len)
    rd = [PDReadDataHolder(*(x.tobytes() for x in r)) for r in reads]
    hp = [PDHaplotypeDataHolder(h[0].tobytes(), h[1].tobytes()) for h in haps]
    cells = sum(len(r[0]) for r in reads) * sum(len(h[0]) for h in haps)
    hmm = IntelPDHMM()
    assert hmm.load()
    hmm.initialize(None)
    out = np.zeros(len(rd) * len(hp))
    hmm.computeLikelihoods(rd, hp, out)  # warm-up
    t0 = time.perf_counter()
    hmm.computeLikelihoods(rd, hp, out)
    e2e_s = time.perf_counter() - t0
    st = hmm.stats()
    ms = hmm.time_runs(a.iters)
    # CPU: GKL's own code on a read sample (flat expansion of the cross product, like pdhmm/JavaData.h:177-242)
    n_cpu = min(a.cpu_reads, a.reads)
    flat = pb.PdhmmBatch.cross(reads[:n_cpu], haps)
    threads = oracle.host_threads()
    res = {"workload": f"config5: {a.reads} reads (len {a.read_len}) x {a.haps} haplotypes (len 200-400) with PD flag bytes",
           "cells": cells, "gpu_kernel_ms": ms, "gpu_kernel_gcups": cells / ms / 1e6,
           "gpu_e2e_gcups_incl_python_marshalling": cells / e2e_s / 1e9,
           "gpu_phases_ms": {"h2d": st.h2d_ms, "kernel": st.kernel_ms, "d2h": st.d2h_ms}, "cpu_threads": threads,
           "cpu_sample": f"first {n_cpu} reads x all haplotypes"}
    if oracle.ref_available():
        for name, level in (("avx512_or_best", 0), ("scalar", 1)):
            oracle.ref_pdhmm(flat.slice(0, 64), level, threads)
            r, rc, secs = oracle.ref_pdhmm(flat, level, threads)
            res[f"cpu_{name}_gcups"] = flat.cells() / secs / 1e9
            res[f"max_abs_diff_vs_cpu_{name}"] = float(np.abs(out[:flat.n] - r).max())
    port = oracle.port_pdhmm(flat, True, threads)[0]
    res["max_abs_diff_vs_serial_restatement"] = float(np.abs(out[:flat.n] - port).max())
    hmm.done()
    print(json.dumps(res))
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(res) + "\n")


if __name__ == "__main__":
    main()
