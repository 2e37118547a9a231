"""Smith-Waterman measurement on the GPU box: pairs per second and cell updates per second of the CUDA path (kernel
resident and end to end through the operator interface) next to GKL's own AVX-512 / AVX2 code on the host cores, with
bit-exact parity over every pair.

    python bench/sw_bench.py [--pairs 14196] [--reps 3] [--out gpurun_out/sw_bench.json]

Workload: the shape of the reference's own test data (src/test/resources/smith-waterman.SOFTCLIP.in: 14 196 pairs,
reference 10-541 bases, median 352, alternate = a mutated copy) regenerated synthetically -- the file itself does not
travel to the GPU box -- aligned with the reference test's parameters (200, -150, -260, -11), SOFTCLIP.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import oracle  # noqa: E402  (checker / CPU baseline only)
from gkl_b200.smithwaterman import IntelSmithWaterman, SWOverhangStrategy, SWParameters, pack  # noqa: E402


def synth_pairs(n, seed=7):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    refs, alts = [], []
    for _ in range(n):
        L = int(np.clip(rng.normal(352, 90), 10, 541))
        x = acgt[rng.integers(0, 4, size=L)]
        y = x.copy()
        for _ in range(int(rng.integers(0, 4))):      # a few SNPs / short indels, like haplotypes against a reference
            pos = int(rng.integers(0, len(y)))
            kind = rng.random()
            if kind < 0.4:
                y[pos] = acgt[rng.integers(0, 4)]
            elif kind < 0.7:
                y = np.delete(y, slice(pos, pos + int(rng.integers(1, 6))))
            else:
                y = np.insert(y, pos, acgt[rng.integers(0, 4, size=int(rng.integers(1, 6)))])
        if len(y) == 0:
            y = x[:1].copy()
        refs.append(x.tobytes())
        alts.append(y.astype(np.uint8).tobytes())
    return refs, alts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=14196)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/sw_bench.json")
    a = ap.parse_args()
    refs, alts = synth_pairs(a.pairs)
    params, strat = (200, -150, -260, -11), 9
    s1, o1 = pack(refs)
    s2, o2 = pack(alts)
    cells = int(np.sum(np.diff(o1) * np.diff(o2)))
    sw = IntelSmithWaterman()
    assert sw.load()
    P = SWParameters(*params)
    cig, off = sw.align_packed(s1, o1, s2, o2, P, strat)  # warm-up
    best = 1e9
    for _ in range(a.reps):
        t0 = time.perf_counter()
        cig, off = sw.align_packed(s1, o1, s2, o2, P, strat)
        best = min(best, time.perf_counter() - t0)
    st = sw.stats()
    best_c = 1e9
    for _ in range(a.reps):  # the C-ABI call alone: H2D, kernel, D2H and the CIGAR strings written by the library
        t0 = time.perf_counter()
        sw.align_packed_raw(s1, o1, s2, o2, P, strat)
        best_c = min(best_c, time.perf_counter() - t0)
    kms = sw.time_runs(a.reps)
    threads = oracle.host_threads()
    res = {"workload": f"{a.pairs} pairs shaped like smith-waterman.SOFTCLIP.in, params {params}, SOFTCLIP",
           "pairs": a.pairs, "cells": cells, "gpu_kernel_ms": kms, "gpu_kernel_gcups": cells / kms / 1e6,
           "gpu_kernel_pairs_per_s": a.pairs / kms * 1e3, "gpu_e2e_ms_c_abi": best_c * 1e3,
           "gpu_e2e_gcups_c_abi": cells / best_c / 1e9, "gpu_e2e_ms_incl_python_strings": best * 1e3, "gpu_phases_ms": {"h2d": st.h2d_ms, "kernel": st.kernel_ms, "d2h": st.d2h_ms},
           "resident_warps": st.warps, "cpu_threads": threads}
    if oracle.ref_available():
        for name, eng in (("avx512_or_best", 0), ("avx2", 1)):
            oracle.ref_sw(s1[:o1[64]], o1[:65], s2[:o2[64]], o2[:65], params, strat, threads)
            rc, ro, secs = oracle.ref_sw(s1, o1, s2, o2, params, strat, threads, engine=eng)
            res[f"cpu_{name}_ms"] = secs * 1e3
            res[f"cpu_{name}_gcups"] = cells / secs / 1e9
            res[f"mismatches_vs_cpu_{name}"] = int(sum(1 for k in range(a.pairs) if rc[k] != cig[k] or ro[k] != off[k]))
        _, _, secs1 = oracle.ref_sw(s1, o1, s2, o2, params, strat, 1)
        res["cpu_1_thread_gcups"] = cells / secs1 / 1e9
    # the reference API as it is: one pair per call (IntelSmithWaterman.align) -- latency, not throughput
    t0 = time.perf_counter()
    for k in range(200):
        sw.align(refs[k], alts[k], P, SWOverhangStrategy.SOFTCLIP)
    res["single_pair_call_us_gpu"] = (time.perf_counter() - t0) / 200 * 1e6
    st1 = sw.stats()
    res["single_pair_last_call_device_us"] = {"h2d": st1.h2d_ms * 1e3, "kernel": st1.kernel_ms * 1e3, "d2h": st1.d2h_ms * 1e3}
    if oracle.ref_available():
        t0 = time.perf_counter()
        for k in range(200):
            a1, b1 = pack([refs[k]])
            a2, b2 = pack([alts[k]])
            oracle.ref_sw(a1, b1, a2, b2, params, strat, 1)
        res["single_pair_call_us_cpu_incl_python"] = (time.perf_counter() - t0) / 200 * 1e6
    pc, po, _ = oracle.port_sw(s1, o1, s2, o2, params, strat, threads)
    res["mismatches_vs_restatement"] = int(sum(1 for k in range(a.pairs) if pc[k] != cig[k] or po[k] != off[k]))
    sw.close()
    print(json.dumps(res))
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(res) + "\n")


if __name__ == "__main__":
    main()
