"""Rerun of the flagged pairs: fp64 list kernel (default) against the range-extended fp32 kernel (GKLB_R2=1), on the
config-2, config-4-shaped and config-3 workloads.  One child process per mode (the switch is read once per process).
    python bench/rerun_compare.py > profiles/r2_rerun_r2_vs_fp64.json
"""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
CHILD = r"""
import json, sys, numpy as np
from gkl_b200 import native, synth
e = native.Engine(0, False)
res = {}
for name, b in (("config2_10000x128", synth.config2()), ("config4_shape_8000x256", synth.config4(8000, 256))):
    e.stage(b); e.run(); e.synchronize()
    ms = e.time_runs(5); st = e.stats()
    res[name] = {"all_kernels_ms": ms, "sweep_ms": st.sweep_ms, "rerun_and_pack_ms": ms - st.sweep_ms,
                 "gcups": b.cells() / ms / 1e6, "flagged_pairs": int(st.fallback_pairs), "fp64_pairs": int(st.fp64_pairs)}
regions = synth.config3(32, seed=3)
cells = sum(r.cells() for r in regions)
t = sw = 0.0
for r in regions:
    e.stage(r); e.run(); e.synchronize(); t += e.time_runs(5); sw += e.stats().sweep_ms
res["config3_per_region"] = {"all_kernels_ms": t, "sweep_ms": sw, "gcups": cells / t / 1e6}
e.stage_multi(regions); e.run(); e.synchronize()
t = e.time_runs(3); st = e.stats()
res["config3_one_job"] = {"all_kernels_ms": t, "sweep_ms": st.sweep_ms, "gcups": cells / t / 1e6,
                          "flagged_pairs": int(st.fallback_pairs), "fp64_pairs": int(st.fp64_pairs)}
print(json.dumps(res))
"""
out = {}
for mode, val in (("fp64_rerun_default", "0"), ("range_extended_fp32_rerun", "1")):
    env = dict(os.environ, GKLB_R2=val, PYTHONPATH=str(ROOT))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, cwd=str(ROOT))
    out[mode] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-800:]}
print(json.dumps(out, indent=1))
