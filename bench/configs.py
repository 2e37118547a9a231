"""The BASELINE.json configurations beside the headline one, measured by bench.py on rank 0 and reported in its
`configs` object.  Each entry carries value / e2e / parity / roofline / cpu_baseline.

  config3  32 HaplotypeCaller-shaped active regions, one synchronous gklb_engine_compute call per region
           (IntelPairHmm.computeLikelihoods is one call per region, IntelPairHmm.java:130-147); also as ONE
           gklb_pairhmm_compute_multi call (the caller-side coalescing of SURVEY.md 8(f) N2)
  config5  PDHMM, 10 000 reads x 128 haplotypes through gklb_pdhmm_compute_cross with host buffers
           (pdhmm/IntelPDHMM.cc:62-137; pdhmm/pdhmm.h:1133-1290)
  config4  1 M reads (len 150) x 256 haplotypes through the product path a JVM uses: one process,
           GKLB_DEVICES=all, gklb_pairhmm_compute shards the reads over the GPUs of the box

The CPU side is GKL's own compiled code (oracle/_ref) on the box's host threads: the parity reference and the
reported baseline.  oracle/ is touched here only as checker / CPU arm.
"""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

FLOP_PER_CELL = 12
PEAKS = {}  # {'fp32': fn, 'fp64': fn} -> (TFLOP/s, source); set by bench.py (bench.py and bench/ share a name)


def _live_dram_bytes(kernel_regex: str, child: list):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch matching kernel_regex, from a child ncu capture (after
    the timed region; ncu replays the kernel, nothing timed runs under it).  None when ncu is not usable."""
    import csv
    import io
    import shutil
    import subprocess
    import sys
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    under_profiler = any(os.environ.get(k) for k in ("CUDA_INJECTION64_PATH", "NV_COMPUTE_PROFILER_PERFWORKS_DIR",
                                                      "NV_NSIGHT_INJECTION_TRANSPORT_TYPE"))
    if not os.path.exists(ncu) or under_profiler or os.environ.get("GKLB_BENCH_NO_NCU"):
        return None
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--print-units", "base",
           "-k", f"regex:{kernel_regex}", "-s", "1", "-c", "1", "--csv", sys.executable, *child]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        hdr = next((row for row in rows if "Metric Name" in row), None)
        if not hdr:
            return None
        ni, vi = hdr.index("Metric Name"), hdr.index("Metric Value")
        vals = [float(row[vi].replace(",", "")) for row in rows[rows.index(hdr) + 1:]
                if len(row) > vi and row[ni].startswith("dram__bytes")]
        return sum(vals) if len(vals) == 2 else None
    except Exception:  # noqa: BLE001
        return None


def _rel(gpu: np.ndarray, cpu: np.ndarray) -> float:
    fin = np.isfinite(cpu)
    if not np.array_equal(np.isfinite(gpu), fin):
        return float("inf")
    return float(np.max(np.abs(gpu[fin] - cpu[fin]) / np.abs(cpu[fin]))) if fin.any() else 0.0


def _cpu_pairhmm(b, threads):
    import oracle
    if oracle.ref_available():
        out, avx512, secs = oracle.ref_pairhmm(b, False, threads=threads)
        return out, secs, "reference", "GKL avx512_impl.cc" if avx512 else "GKL avx_impl.cc"
    out, _, secs = oracle.port_pairhmm(b, False, threads=threads)
    return out, secs, "port", "oracle/pairhmm_oracle.c"


def config3(device: int = 0, seed: int = 3) -> dict:
    import oracle
    from gkl_b200 import native, synth
    regions = synth.config3(32, seed=seed)
    cells = sum(r.cells() for r in regions)
    pairs = sum(r.n_reads * r.n_haps for r in regions)
    eng = native.Engine(device, False)
    outs = [np.empty(r.n_reads * r.n_haps) for r in regions]
    for r, o in zip(regions, outs):  # warm-up (allocations, first launches)
        eng.compute(r, out=o)
    # resident: each region staged in HBM, its kernels timed with CUDA events (gklb_engine_time_runs)
    kernel_ms = sweep_ms = 0.0
    launches, fallback, classes = 0, 0, 0
    names = set()
    for r in regions:
        eng.stage(r)
        eng.run()
        st = eng.stats()
        sweep_ms += st.sweep_ms
        launches += st.kernel_launches
        classes = max(classes, st.n_classes)
        names.add(eng.sweep_kernel().split(" (")[0])
        kernel_ms += eng.time_runs(5)
    # end to end: 32 sequential synchronous calls from pageable host buffers, best of 5 rounds
    best = 1e30
    for _ in range(5):
        t0 = time.perf_counter()
        for r, o in zip(regions, outs):
            eng.compute(r, out=o)
        best = min(best, time.perf_counter() - t0)
        fallback = 0
    for r, o in zip(regions, outs):
        eng.compute(r, out=o)
        fallback += int(eng.stats().fallback_pairs)
    eng.close()
    res = {
        "workload": "configs[2]: 32 active regions x ~300 reads (len 35-250) x ~64 haplotypes, one synchronous call per region",
        "cells": cells, "pairs": pairs, "value": cells / kernel_ms / 1e6, "unit": "GCUPS",
        "value_note": "kernels only, each region staged in HBM, CUDA events (5 runs per region)",
        "e2e": {"value": cells / best / 1e9, "unit": "GCUPS", "ms_per_call": best * 1e3 / len(regions),
                "api": "32 x gklb_engine_compute, pageable host buffers, best of 5 rounds",
                "h2d_bytes": int(sum(r.input_bytes() + 8 * (r.n_reads + r.n_haps + 2) for r in regions)),
                "d2h_bytes": 8 * pairs},
        "gpu_launches": launches, "classes_per_call_max": classes, "fallback_pairs": fallback,
    }
    multi = getattr(native, "global_compute_multi", None)
    if multi is not None:
        try:
            native.global_init(False, 1)
            outs_m = multi(regions)
            best_m = 1e30
            for _ in range(5):
                t0 = time.perf_counter()
                outs_m = multi(regions, outs_m)
                best_m = min(best_m, time.perf_counter() - t0)
            res["e2e_multi"] = {"value": cells / best_m / 1e9, "unit": "GCUPS", "ms_total": best_m * 1e3,
                                "api": "ONE gklb_pairhmm_compute_multi call over the 32 regions, pageable host buffers",
                                "identical_to_per_region_calls": bool(all(np.array_equal(a, c) for a, c in zip(outs, outs_m)))}
            native.global_done()
        except Exception as ex:
            res["e2e_multi"] = {"error": f"{type(ex).__name__}: {ex}"}
    threads = oracle.host_threads()
    t_cpu, err = 0.0, 0.0
    kind = detail = ""
    for r, o in zip(regions, outs):
        ref, secs, kind, detail = _cpu_pairhmm(r, threads)
        t_cpu += secs
        err = max(err, _rel(o, ref))
    peak, peak_src = PEAKS['fp32']()
    ach = cells * FLOP_PER_CELL / (sweep_ms * 1e-3) / 1e12
    res["parity"] = {"max_rel_err": err, "pairs_checked": pairs, "against": f"{kind} ({detail}), every pair"}
    res["roofline"] = {"bound": "fp32", "kernel": " + ".join(sorted(names)), "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                       "frac": ach / peak, "peak_source": peak_src, "traffic": None,
                       "note": "sweep launches of the 32 regions summed (CUDA events); small launches fill the GPU only partly"}
    res["cpu_baseline"] = {"value": cells / t_cpu / 1e9, "unit": "GCUPS", "cores": threads, "kind": kind,
                           "sample": f"all 32 regions once ({detail}, pair loop only, {t_cpu:.2f} s)"}
    return res


def config5(device: int = 0) -> dict:
    import oracle
    from gkl_b200 import synth
    from gkl_b200.pdhmm import IntelPDHMM
    from gkl_b200.pdhmm_batch import PdhmmBatch
    os.environ.setdefault("GKLB_DEVICE", str(device))
    reads, haps = synth.config5()
    ops = PdhmmBatch.operands(reads, haps)
    R, H = len(reads), len(haps)
    cells = int(ops.read_lengths.sum()) * int(ops.hap_lengths.sum())
    hmm = IntelPDHMM()
    if not hmm.load():
        raise RuntimeError("no sm_100 device")
    hmm.initialize(None)
    out = np.zeros(R * H)
    hmm.compute_cross(ops, out)  # warm-up
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        hmm.compute_cross(ops, out)
        best = min(best, time.perf_counter() - t0)
    st = hmm.stats()
    kernel_ms = hmm.time_runs(3)
    kernel_name = hmm.kernel_name()
    hmm.done()
    threads = oracle.host_threads()
    # parity: every pair against the reference's scalar path (what GATK's Java produced the golden files with);
    # the restatement is bit-identical to GKL's pdhmm-serial.cc (tests/test_oracle_pdhmm.py) and runs on all threads
    err, t_chk = 0.0, 0.0
    step = 500
    for r0 in range(0, R, step):
        r1 = min(R, r0 + step)
        flat = ops.expand_cross(r0, r1)
        ref, rc, secs = oracle.port_pdhmm(flat, True, threads)
        t_chk += secs
        err = max(err, float(np.max(np.abs(out[r0 * H:r1 * H] - ref))))
    # CPU baseline: GKL's own fastest PDHMM (AVX-512 where the host has it) on a bounded sample
    n_cpu = min(1000, R)
    sample = ops.expand_cross(0, n_cpu)
    kind, detail, cpu_gcups, dev_avx = "port", "oracle/pdhmm_oracle.c", 0.0, None
    if oracle.ref_available():
        oracle.ref_pdhmm(sample.slice(0, 512), 0, threads)
        ref_fast, rc, secs = oracle.ref_pdhmm(sample, 0, threads)
        cpu_gcups = sample.cells() / secs / 1e9
        kind, detail = "reference", "GKL computePDHMM, fastest available (AVX-512 / AVX2) + OpenMP"
        dev_avx = float(np.max(np.abs(out[:sample.n] - ref_fast)))
    else:
        _, _, secs = oracle.port_pdhmm(sample, True, threads)
        cpu_gcups = sample.cells() / secs / 1e9
    peak, peak_src = PEAKS['fp64']()
    ach = cells * FLOP_PER_CELL / (kernel_ms * 1e-3) / 1e12
    traffic = _live_dram_bytes("k_pdhmm3", [str(Path(__file__).resolve().parent / "profile_c5.py"), str(R)])
    return {
        "workload": "configs[4]: PDHMM, 10000 reads (len 101) x 128 haplotypes (len 200-400) with PD flag bytes",
        "cells": cells, "pairs": R * H, "value": cells / kernel_ms / 1e6, "unit": "GCUPS", "dtype": "f64",
        "value_note": "kernel only, operands resident in HBM, CUDA events (gklb_pdhmm_time_runs, 3 runs)",
        "e2e": {"value": cells / best / 1e9, "unit": "GCUPS", "ms": best * 1e3,
                "phases_ms": {"h2d": st.h2d_ms, "kernel": st.kernel_ms, "d2h": st.d2h_ms},
                "api": "gklb_pdhmm_compute_cross (what IntelPDHMM.computeLikelihoodsNative calls), host buffers, best of 3",
                "h2d_bytes": int(2 * H * ops.max_hap + 5 * R * ops.max_read + 8 * (R + H)), "d2h_bytes": 8 * R * H},
        "gpu_launches": int(st.kernel_launches),
        "parity": {"max_abs_err": err, "pairs_checked": R * H,
                   "against": f"restatement of pdhmm-serial.cc (bit-identical to GKL's scalar path), every pair, {t_chk:.1f} s",
                   "max_abs_diff_vs_gkl_fastest_on_cpu_sample": dev_avx, "tolerance": "1e-4 absolute (IntelPDHMMUnitTest.java:33)"},
        "roofline": {"bound": "fp64", "kernel": kernel_name, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                     "frac": ach / peak, "peak_source": peak_src, "flop_per_cell": FLOP_PER_CELL,
                     "traffic": traffic, "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, child "
                     "ncu capture after the timed region (null: ncu not usable); algorithmic bytes "
                     f"{int(2 * H * ops.max_hap + 5 * R * ops.max_read + 8 * R * H)}"},
        "cpu_baseline": {"value": cpu_gcups, "unit": "GCUPS", "cores": threads, "kind": kind,
                         "sample": f"first {n_cpu} reads x all haplotypes ({detail})"},
    }


def config4(n_gpus: int, reads: int = 1_000_000, haps: int = 256) -> dict:
    import oracle
    from gkl_b200 import native, synth
    t0 = time.time()
    b = synth.config4(reads, haps)
    gen_s = time.time() - t0
    cells = b.cells()
    os.environ["GKLB_DEVICES"] = "all"
    res = {"workload": f"configs[3]: {reads} reads (len 150) x {haps} haplotypes (len 200-400), one gklb_pairhmm_compute call, "
                       f"reads sharded over {n_gpus} GPUs inside the process (GKLB_DEVICES=all)",
           "cells": cells, "pairs": reads * haps, "generate_s": gen_s}
    out = np.empty(reads * haps, dtype=np.float64)
    modes = {}
    for mode in ("direct", "nccl"):
        os.environ["GKLB_SHARD"] = mode
        try:
            n_dev = native.global_init(False, 1)
            native.global_compute(b, out)  # warm-up: engine buffers, first launches
            best = 1e30
            for _ in range(2):
                t0 = time.perf_counter()
                native.global_compute(b, out)
                best = min(best, time.perf_counter() - t0)
            st = native.global_stats()
            modes[mode] = {"devices": n_dev, "e2e_s": best, "e2e_gcups": cells / best / 1e9,
                           "phases_ms_max_over_devices": {"h2d_pack": st.h2d_ms, "kernels": st.kernel_ms, "d2h": st.d2h_ms},
                           "kernels_gcups": cells / (st.kernel_ms * 1e-3) / 1e9, "fallback_pairs": int(st.fallback_pairs),
                           "launches": int(st.kernel_launches)}
            native.global_done()
        except Exception as ex:
            modes[mode] = {"error": f"{type(ex).__name__}: {ex}"}
            try:
                native.global_done()
            except Exception:
                pass
    ok = {k: v for k, v in modes.items() if "error" not in v}
    if not ok:
        res["error"] = modes
        return res
    fastest = min(ok, key=lambda k: ok[k]["e2e_s"])
    # leave `out` from the default (direct) mode if it ran, else from whatever ran last
    os.environ["GKLB_SHARD"] = "direct" if "direct" in ok else fastest
    native.global_init(False, 1)
    native.global_compute(b, out)
    native.global_done()
    m = ok["direct" if "direct" in ok else fastest]
    res["value"] = m["kernels_gcups"]
    res["unit"] = "GCUPS"
    res["value_note"] = ("cells / slowest device's kernel phase (CUDA events) inside the sharded call; direct mode: a device's "
                         "shard runs as up to 12 pieces on two engines in turn, the kernel phase is the span from the first "
                         "piece's first kernel to the last piece's last kernel, and the h2d/d2h figures are the first piece's way in "
                         "and the last piece's way out (the other copies run under the kernels of the neighbouring piece)")
    res["e2e"] = {"value": m["e2e_gcups"], "unit": "GCUPS", "seconds": m["e2e_s"],
                  "api": "gklb_pairhmm_compute, pageable host buffers, best of 2 after warm-up",
                  "h2d_bytes": int(b.input_bytes() * 1 + 8 * (reads + haps + 2)), "d2h_bytes": 8 * reads * haps}
    res["sharding_modes"] = modes
    res["default_mode"] = "direct"
    res["fastest_mode"] = fastest
    res["gpu_launches"] = m["launches"]
    # parity: a 1 % stratified read sample (every 100th read, all haplotypes) + every pair that took the fp64 rerun
    threads = oracle.host_threads()
    idx = np.arange(0, reads, 100)
    sample = synth.PairHmmBatch(
        np.concatenate([[0], np.cumsum(b.read_lens[idx])]).astype(np.int64),
        *[np.concatenate([a[b.read_off[i]:b.read_off[i + 1]] for i in idx]) for a in
          (b.read_bases, b.read_quals, b.ins_gop, b.del_gop, b.gcp)], b.hap_off, b.hap_bases)
    ref, secs, kind, detail = _cpu_pairhmm(sample, threads)
    got = out.reshape(reads, haps)[idx].ravel()
    err_sample = _rel(got, ref)
    # pairs below GKL's threshold: log10(1e-28) - log10(2^120) = -64.12; a small margin catches the borderline ones
    fb = np.flatnonzero(out < -64.0)
    n_fb_all = int(len(fb))
    cap = int(os.environ.get("GKLB_C4_MAX_RERUN_CHECK", "0"))  # development runs may bound the CPU work; 0 = all pairs
    if cap and len(fb) > cap:
        fb = fb[:: len(fb) // cap + 1]
    err_fb, fb_checked, fb_secs = None, 0, 0.0
    if oracle.ref_available() and len(fb):
        pr, ph = (fb // haps).astype(np.int32), (fb % haps).astype(np.int32)
        ref_fb, fb_secs = oracle.ref_pairhmm_pairs(b, pr, ph, threads)
        err_fb = _rel(out[fb], ref_fb)
        fb_checked = int(len(fb))
    res["parity"] = {"max_rel_err_sample": err_sample, "sample": f"every 100th read x all haplotypes ({len(idx) * haps} pairs)",
                     "max_rel_err_all_rerun_pairs": err_fb, "rerun_pairs_checked": fb_checked,
                     "pairs_below_threshold": n_fb_all,
                     "rerun_pairs_reported_by_engine": m["fallback_pairs"], "rerun_check_seconds": fb_secs,
                     "against": f"{kind} ({detail})"}
    peak, peak_src = PEAKS['fp32']()
    ach = cells * FLOP_PER_CELL / (m["phases_ms_max_over_devices"]["kernels"] * 1e-3) / 1e12 / n_gpus
    probe = native.Engine(0, False)   # the sweep kernel the engine picks for this shape (a shard plans the same classes)
    probe.stage(synth.config4(2000, haps))
    sweep_name = probe.sweep_kernel()
    probe.close()
    res["roofline"] = {"bound": "fp32", "kernel": f"{sweep_name} + k_sweep_list<VD1> (all kernels of the slowest device)",
                       "achieved": ach, "peak": peak, "unit": "TFLOP/s per GPU", "frac": ach / peak, "peak_source": peak_src,
                       "traffic": None}
    res["cpu_baseline"] = {"value": sample.cells() / secs / 1e9, "unit": "GCUPS", "cores": threads, "kind": kind,
                           "sample": f"1 % of the reads (every 100th) x all haplotypes, {secs:.2f} s; EXTRAPOLATION to the "
                                     f"full batch: {cells / (sample.cells() / secs):.0f} s"}
    return res
