#!/usr/bin/env python
"""bench.py -- PairHMM cell-updates/sec (GCUPS) of the B200 engine, next to GKL's AVX-512+OpenMP code.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (IntelPairHmm.computeLikelihoodsNative's pair loop) over one
synthetic read x haplotype batch.  N=1 runs BASELINE.json configs[1]: 10 000 reads (len 101) x 128
haplotypes (len 200-400), empirical quality distribution (gkl_b200.synth.config2).  N>1 shards reads
(weak scaling: every rank owns 10 000 reads of the same shape against the same 128 haplotypes), with an
NCCL broadcast of the haplotype panel and an NCCL gather of the likelihood slabs inside the timed step.

Printed keys (one JSON line, rank 0):
  value        GCUPS with the batch already resident in HBM (CUDA events around K steps, L2 flushed
               between steps, max over ranks)
  e2e          the same metric through the C-ABI with HOST buffers: per step the packed arenas are copied
               host->device, kernels run, likelihoods come back device->host
  roofline     the forward-sweep kernel against the measured fp32 FMA rate of this GPU (the recurrence is
               CUDA-core fp32 bound: 12 flop and ~4e-4 HBM bytes per cell), with the HBM view beside it
  cpu_baseline GKL's own AVX-512/AVX PairHMM (oracle/_ref) or the oracle port, all host threads
  configs      the other BASELINE.json configurations, each with value / e2e / parity / roofline / cpu_baseline:
               c3 (32 HaplotypeCaller-shaped regions) and c5 (PDHMM) at N=1, c4 (1 M x 256 through the
               in-process multi-GPU product path) at N=8 (bench/configs.py)
  --impl reference  times only the CPU arm and prints it in the same shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "pairhmm_cell_updates_per_sec"
UNIT = "GCUPS"
FLOP_PER_CELL = 12  # 8 mul + 4 add, avx-pairhmm-template.h:213-222


def env_int(name, default):
    return int(os.environ.get(name, default))


def workload_name(n_reads: int, n_haps: int) -> str:
    """One string for both arms (the driver compares them)."""
    return (f"configs[1]: {n_reads} reads (len 101) x {n_haps} haplotypes (len 200-400) per GPU, empirical quals; "
            "fp32 forward sweep + fp64 rerun of pairs under 1e-28 (IntelPairHmm.cc:150-169)")


def workload(rank: int, n_reads: int, n_haps: int):
    from gkl_b200 import synth
    b = synth.config2(n_reads, n_haps, 101, seed=2)
    if rank == 0:
        return b
    # other ranks: a different read shard against the same panel (same seed for the panel, different reads)
    rng = np.random.default_rng(1000 + rank)
    haps = [b.hap_bases[b.hap_off[h]:b.hap_off[h + 1]] for h in range(b.n_haps)]
    reads = synth._reads_from_panel(rng, haps, np.full(n_reads, 101, dtype=np.int64), empirical=True)
    return synth._assemble(haps, reads)


def algorithmic_bytes(b) -> int:
    return 5 * int(b.read_off[-1]) + int(b.hap_off[-1]) + 8 * b.n_reads * b.n_haps


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons = index, threading.Event(), [], set()
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self.stop_flag.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as ex:  # clocks are evidence, not a dependency
            self.reasons.add(f"unavailable: {type(ex).__name__}")

    def result(self):
        self.stop_flag.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def _micro_peak(exe_name: str, fallback_file: str, nominal: float, nominal_src: str):
    exe = ROOT / "bench" / "micro" / exe_name
    try:
        out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60).stdout
        rows = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
        best = max(r["tflops"] for r in rows if "tflops" in r)
        return best, f"measured live: bench/micro/{exe_name} (best variant)"
    except Exception:
        pass
    try:
        rows = [json.loads(l) for l in (ROOT / "profiles" / fallback_file).read_text().splitlines()]
        return max(r["tflops"] for r in rows if "tflops" in r), f"profiles/{fallback_file} (measured on this pool)"
    except Exception:
        return nominal, nominal_src


def fp32_peak_tflops():
    """Measured fp32 FMA peak of this GPU (bench/micro/fp32_peak, built by __graft_entry__.build)."""
    return _micro_peak("fp32_peak", "r2_fp32_peak.jsonl", 148 * 128 * 2 * 1.965e9 / 1e12,
                       "nominal 148 SM x 128 lanes x 2 x 1.965 GHz")


def fp64_peak_tflops():
    """Measured fp64 FMA peak of this GPU (bench/micro/fp64_peak)."""
    return _micro_peak("fp64_peak", "r2_fp64_peak.jsonl", 148 * 64 * 2 * 1.965e9 / 1e12,
                       "nominal 148 SM x 64 lanes x 2 x 1.965 GHz")


def measured_hbm_gbs():
    try:
        return json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"], "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def dram_traffic_record(reads: int, haps: int, live: bool = True):
    """DRAM bytes per launch of the dominant kernel.  Measured live when ncu is usable on this box: one launch of the
    sweep on the same workload in a child process, after the timed region (ncu replays the kernel, so nothing timed
    runs under it).  Falls back to the committed capture (profiles/r2_traffic.json, which names its commit)."""
    import csv
    import io
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    under_profiler = any(os.environ.get(k) for k in ("CUDA_INJECTION64_PATH", "NV_COMPUTE_PROFILER_PERFWORKS_DIR",
                                                      "NV_NSIGHT_INJECTION_TRANSPORT_TYPE"))
    if live and os.path.exists(ncu) and not under_profiler and not os.environ.get("GKLB_BENCH_NO_NCU"):
        cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--print-units",
               "base", "-k", "regex:k_h2|k_sweep_tasks", "-s", "2", "-c", "1", "--csv", sys.executable,
               str(ROOT / "bench" / "profile_c2.py"), str(reads), "101", str(haps)]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=str(ROOT))
            vals, kernel = {}, None
            rows = list(csv.reader(io.StringIO(r.stdout)))
            hdr = next((row for row in rows if "Metric Name" in row), None)
            if hdr:
                ni, vi, ki = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Kernel Name")
                for row in rows[rows.index(hdr) + 1:]:
                    if len(row) > vi and row[ni].startswith("dram__bytes"):
                        vals[row[ni]] = float(row[vi].replace(",", ""))
                        kernel = row[ki]
            if len(vals) == 2:
                total = vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
                return {"kernel": kernel, "workload": f"{reads} x {haps}, one launch, measured by this run",
                        "dram_bytes_read": vals["dram__bytes_read.sum"], "dram_bytes_write": vals["dram__bytes_write.sum"],
                        "dram_bytes_per_launch": total, "how": " ".join(cmd[:13]) + " python bench/profile_c2.py ...",
                        "note": "the likelihoods stay in the 126 MB L2 until the device->host copy, so DRAM sees mostly "
                                "the input records"}
        except Exception:  # noqa: BLE001 -- no counters permission, ncu missing, timeout: use the committed capture
            pass
    try:
        rec = json.loads((ROOT / "profiles" / "r2_traffic.json").read_text())
        rec["note"] = "committed capture (ncu was not usable in this run); " + rec.get("note", "")
        return rec
    except Exception:
        return None


def cpu_arm(b, steps: int, warmup: int):
    """GKL's own compiled PairHMM (or the oracle port) on all host threads; returns (gcups, info).
    info["out"] holds the CPU likelihoods of the batch (the parity reference of the same run).  The figure is the
    best of `steps` timed passes after warm-up."""
    import oracle
    threads = oracle.host_threads()
    if oracle.ref_available():
        fn = lambda: oracle.ref_pairhmm(b, False, threads=threads)
        out, avx512, _ = fn()
        kind, detail = "reference", "GKL avx512_impl.cc" if avx512 else "GKL avx_impl.cc"
    else:
        fn = lambda: oracle.port_pairhmm(b, False, threads=threads)
        out = fn()[0]
        kind, detail = "port", "oracle/pairhmm_oracle.c"
    for _ in range(max(0, warmup - 1)):
        fn()
    secs = [fn()[2] for _ in range(steps)]
    best = min(secs)
    return b.cells() / best / 1e9, {"kind": kind, "cores": threads, "detail": detail, "seconds_per_step": best,
                                    "seconds_mean": sum(secs) / len(secs), "out": out}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    b = workload(0, args.reads, args.haps)
    t0 = time.time()
    gcups, info = cpu_arm(b, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": gcups, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": info["seconds_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(b.n_reads, b.n_haps)},
        "cells_per_step": b.cells(),
        "cpu_baseline": {"value": gcups, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                         "sample": f"the full batch, best of {args.steps} passes ({info['detail']}, OpenMP "
                                   f"schedule(dynamic,1), pair loop only; mean {info['seconds_mean'] * 1e3:.1f} ms); "
                                   f"wall {time.time() - t0:.1f} s"},
        "e2e": {"value": gcups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class _DevArray:
    """__cuda_array_interface__ view of an engine-owned device buffer (so torch can gather it)."""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def parity_all_pairs(gpu: np.ndarray, cpu: np.ndarray) -> float:
    """max_i |gpu - cpu| / |cpu| over every pair; a non-finite GPU value where the CPU value is finite (or the other
    way round) is a failure, reported as inf."""
    fin = np.isfinite(cpu)
    if not np.array_equal(np.isfinite(gpu), fin):
        return float("inf")
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(gpu[fin] - cpu[fin]) / np.abs(cpu[fin])))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from gkl_b200 import native

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    torch.cuda.set_device(local)
    distributed = world > 1
    dev = torch.device("cuda", local)
    cpu_group = None
    if distributed:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")  # host-side barriers that do not occupy the GPUs

    b = workload(rank, args.reads, args.haps)
    cells = b.cells()
    pairs = b.n_reads * b.n_haps
    eng = native.Engine(local, False)
    stream = torch.cuda.Stream(device=dev)  # one stream for the engine's kernels, the collectives and the events
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)

    # ---- resident path: inputs staged in HBM, K steps timed with CUDA events ----
    hap_dev = torch.from_numpy(b.hap_bases).to(dev)
    arenas_dev = [torch.from_numpy(x).to(dev) for x in (b.read_bases, b.read_quals, b.ins_gop, b.del_gop, b.gcp)]
    eng.stage(b, arenas=arenas_dev, hap=hap_dev, device=True)
    sweep_kernel = eng.sweep_kernel()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    panel_buf = hap_dev if rank == 0 else torch.empty_like(hap_dev)
    gather_bufs, cap, packed_bytes = None, 0, 0
    if distributed:
        # capacity of the override list (pairs that took the fp64 rerun), agreed once outside the timed region
        eng.run()
        eng.fetch(pairs)
        capt = torch.tensor([int(eng.stats().fallback_pairs)], dtype=torch.int64, device=dev)
        dist.all_reduce(capt, op=dist.ReduceOp.MAX)
        cap = (int(capt[0]) * 5 // 4 + 4096) // 1024 * 1024
        _, packed_bytes = eng.narrow(cap)
        if rank == 0:
            gather_bufs = [torch.empty(packed_bytes, dtype=torch.uint8, device=dev) for _ in range(world)]

    def step():
        if distributed:
            # the haplotype panel travels from GPU 0 over NVLink (one broadcast of the packed bases; the lengths are
            # part of the staged batch) and is consumed by this step's sweep; the results go back to GPU 0 as what
            # the reference's values really are: an fp32 per pair + (index, fp64) overrides for the rerun pairs
            dist.broadcast(panel_buf, src=0)
            eng.update_haps_device(panel_buf)
        eng.run()
        if distributed:
            ptr, nbytes = eng.narrow(cap)
            dist.gather(torch.as_tensor(_DevArray(ptr, nbytes, "|u1"), device=dev), gather_bufs, dst=0)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sweep_ms, launches = 0.0, 0
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record(stream)
        step()
        ev[i][1].record(stream)
        st = eng.stats()  # synchronises the stream; reads the sweep-kernel events of this step
        sweep_ms += st.sweep_ms
        launches += st.kernel_launches
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    dev_ms = sum(a.elapsed_time(c) for a, c in ev)
    clocks = sampler.result()
    resident_out = eng.fetch(pairs)
    fallback = int(eng.stats().fallback_pairs)

    # ---- end to end: host buffers in, host likelihoods out, through the C-ABI compute call ----
    pin = lambda a: torch.from_numpy(a).pin_memory()
    host_arenas = [pin(x) for x in (b.read_bases, b.read_quals, b.ins_gop, b.del_gop, b.gcp)]
    host_hap = pin(b.hap_bases)
    host_out = torch.empty(pairs, dtype=torch.float64).pin_memory()
    e2e_steps = max(3, args.steps)
    for _ in range(2):
        eng.compute(b, arenas=host_arenas, hap=host_hap, out_ptr=host_out.data_ptr())
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.compute(b, arenas=host_arenas, hap=host_hap, out_ptr=host_out.data_ptr())
    e2e_s = time.perf_counter() - t0
    e2e_stats = eng.stats()
    # the same with pageable buffers (what a JNI caller's std::vector arenas and a JVM-pinned double[] look like)
    page_out = np.empty(pairs, dtype=np.float64)
    eng.compute(b, out=page_out)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.compute(b, out=page_out)
    e2e_page_s = time.perf_counter() - t0

    times = torch.tensor([dev_ms, e2e_s * 1e3, float(cells), e2e_page_s * 1e3], dtype=torch.float64, device=dev)
    if distributed:
        tmax = times.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = times.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, e2e_ms_total, total_cells, e2e_page_ms = float(tmax[0]), float(tmax[1]), float(tsum[2]), float(tmax[3])
    else:
        e2e_ms_total, total_cells, e2e_page_ms = e2e_s * 1e3, float(cells), e2e_page_s * 1e3
    torch.cuda.synchronize()
    eng.close()
    del flush, arenas_dev
    torch.cuda.empty_cache()

    line = None
    if rank == 0:
        value = total_cells * args.steps / (dev_ms * 1e-3) / 1e9
        e2e_value = total_cells * e2e_steps / (e2e_ms_total * 1e-3) / 1e9
        peak_tf, peak_src = fp32_peak_tflops()
        hbm, hbm_src = measured_hbm_gbs()
        sweep_per_launch_ms = sweep_ms / max(1, args.steps)
        ach_tf = cells * FLOP_PER_CELL / (sweep_per_launch_ms * 1e-3) / 1e12
        ach_gbs = algorithmic_bytes(b) / (sweep_per_launch_ms * 1e-3) / 1e9
        traffic = dram_traffic_record(args.reads, args.haps, live=(world == 1))
        cpu_gcups, cpu_info = cpu_arm(b, 1, 1)
        cpu_out = cpu_info["out"]
        import oracle  # the CPU arm: one more figure, a single host thread on a 300-read slice of the same batch
        one = b.read_slice(0, min(300, b.n_reads))
        fn1 = oracle.ref_pairhmm if oracle.ref_available() else oracle.port_pairhmm
        fn1(one, False, threads=1)
        cpu_1t = one.cells() / fn1(one, False, threads=1)[2] / 1e9
        parity = max(parity_all_pairs(resident_out, cpu_out), parity_all_pairs(host_out.numpy(), cpu_out),
                     parity_all_pairs(page_out, cpu_out))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(b.n_reads, b.n_haps)},
            "cells_per_step": total_cells, "pairs_per_gpu": pairs, "fallback_pairs_rank0": fallback,
            "l2": "flushed between steps (256 MiB write)",
            "parallelism": (f"reads sharded over {world} GPUs, one process per GPU; per step one NCCL broadcast of the "
                            "haplotype panel and one NCCL gather of the results (fp32 matrix + fp64 overrides of the rerun "
                            f"pairs, {packed_bytes} bytes per GPU) to GPU 0") if world > 1 else "single GPU",
            "parity_max_rel_err_all_pairs_vs_cpu_baseline": parity,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(5 * int(b.read_off[-1]) + int(b.hap_off[-1]) + 8 * (b.n_reads + 1)
                                              + 8 * (b.n_haps + 1)) * world,
                    "d2h_bytes_per_step": 8 * pairs * world, "ms_per_step": e2e_ms_total / e2e_steps,
                    "phases_ms_rank0": {"h2d_pack": e2e_stats.h2d_ms, "kernels": e2e_stats.kernel_ms,
                                        "d2h": e2e_stats.d2h_ms},
                    "api": "gklb_engine_compute (what computeLikelihoodsNative calls), pinned host buffers",
                    "pageable_host_buffers": {"value": total_cells * e2e_steps / (e2e_page_ms * 1e-3) / 1e9,
                                              "ms_per_step": e2e_page_ms / e2e_steps}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "fp32", "kernel": sweep_kernel, "achieved": ach_tf, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                         "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                         "traffic_source": traffic,
                         "flop_per_cell": FLOP_PER_CELL, "ms_per_launch": sweep_per_launch_ms, "peak_source": peak_src,
                         "hbm": {"achieved": ach_gbs, "peak": hbm, "unit": "GB/s", "frac": ach_gbs / hbm,
                                 "peak_source": hbm_src, "algorithmic_bytes_per_launch": algorithmic_bytes(b)},
                         "note": "scalar fp32 recurrence on CUDA cores: neither HBM nor tensor bound"},
            "cpu_baseline": {"value": cpu_gcups, "unit": UNIT, "cores": cpu_info["cores"], "kind": cpu_info["kind"],
                             "value_1_thread": cpu_1t,
                             "sample": f"the full rank-0 batch once ({cpu_info['detail']}, all host threads, pair loop only, "
                                       f"{cpu_info['seconds_per_step']:.2f} s)"},
        }

    # ---- the other BASELINE configurations (rank 0; the other ranks idle on a host-side barrier) ----
    if not args.no_configs:
        import importlib.util
        spec = importlib.util.spec_from_file_location("gklb_bench_configs", ROOT / "bench" / "configs.py")
        extra = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(extra)
        extra.PEAKS.update(fp32=fp32_peak_tflops, fp64=fp64_peak_tflops)
        if rank == 0:
            cfgs = {}
            try:
                if world == 1:
                    cfgs["c3"] = extra.config3(local)
                    cfgs["c5"] = extra.config5(local)
                elif world == 8 or args.force_c4:
                    cfgs["c4"] = extra.config4(world, reads=args.c4_reads)
            except Exception as ex:  # the headline line must survive a failure here
                cfgs["error"] = f"{type(ex).__name__}: {ex}"
            line["configs"] = cfgs
        if distributed:
            dist.barrier(group=cpu_group)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000, help="reads per GPU (BASELINE configs[1]: 10000)")
    ap.add_argument("--haps", type=int, default=128)
    ap.add_argument("--no-configs", action="store_true", help="skip the configs object (c3/c5 at N=1, c4 at N=8)")
    ap.add_argument("--force-c4", action="store_true", help="run config 4 at any N > 1 (measurement)")
    ap.add_argument("--c4-reads", type=int, default=1_000_000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        args.steps = min(args.steps, 20)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
