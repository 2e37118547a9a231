"""Summarise an .ncu-rep (read here, no GPU needed): key throughput metrics, stall mix, hottest SASS lines.

    python bench/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25] > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        print(f"== kernel: {d.get('Kernel Name', '?')}  id={d.get('ID', '?')}")
        for k in KEYS:
            if k in d:
                print(f"  {k:78s} {d[k]:>16s} {units[hdr.index(k)]}")
        stalls = {k: float(v) for k, v in d.items()
                  if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
        print("  stall reasons (warps stalled per issue-active cycle):")
        for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]:
            print(f"    {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:24s} {v:.3f}")
    src = page(rep, "source")
    h = src[1]
    ix = {n: i for i, n in enumerate(h)}
    rows = []
    for r in src[2:]:
        try:
            rows.append((int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]), r[ix["Source"]], r))
        except (ValueError, IndexError):
            pass
    total = sum(r[0] for r in rows)
    print(f"== hottest SASS lines by stall samples (total samples {total})")
    cols = ["stall_wait", "stall_dispatch", "stall_math", "stall_short_sb", "stall_long_sb", "stall_not_selected",
            "stall_branch_resolving"]
    print(f"  {'samples':>8s} {'executed':>10s}  {'wait disp math ssb lsb notsel branch':36s} instruction")
    for smp, ex, text, r in sorted(rows, key=lambda x: -x[0])[:top]:
        print(f"  {smp:8d} {ex:10d}  {' '.join(r[ix[c]] for c in cols):36s} {text[:90]}")
    # opcode mix weighted by execution count
    mix = {}
    for smp, ex, text, r in rows:
        op = text.split()[0] if not text.startswith("@") else text.split()[1]
        op = op.split(".")[0]
        mix[op] = mix.get(op, 0) + ex
    tot = sum(mix.values())
    print("== executed warp-instruction mix")
    for op, n in sorted(mix.items(), key=lambda kv: -kv[1])[:14]:
        print(f"  {op:10s} {n:12d} {100.0 * n / tot:5.1f}%")


if __name__ == "__main__":
    main()
