"""GPU: the range-extended fp32 rerun (GKLB_R2=1, gkl_b200/csrc/pairhmm_r2.cuh) against the reference.

The switch is read once per process, so the checks run in a child process with the variable set."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

CHILD = r"""
import numpy as np, oracle
from gkl_b200 import native, synth
def checker(b):
    t = oracle.host_threads()
    return (oracle.ref_pairhmm(b, False, threads=t) if oracle.ref_available() else oracle.port_pairhmm(b, False, threads=t))[0]
e = native.Engine(0, False)
tot_fb = tot_64 = 0
cases = [synth.random_batch(31, 120, 30), synth.random_batch(32, 120, 30, low_quality=0.1, unrelated=0.3),
         synth.random_batch(35, 200, 40, unrelated=0.6), synth.random_batch(36, 60, 20, read_len=(1, 40), hap_len=(1, 60)),
         synth.config2(1500, 128), synth.config4(600, 256)] + synth.config3(3, seed=5)
for i, b in enumerate(cases):
    out, ref = e.compute(b), checker(b)
    ok = np.isfinite(ref)
    assert np.array_equal(np.isfinite(out), ok), i
    err = float(np.max(np.abs(out[ok] - ref[ok]) / np.abs(ref[ok])))
    st = e.stats()
    tot_fb += st.fallback_pairs; tot_64 += st.fp64_pairs
    assert err <= 1e-5, (i, err)
    assert st.fp64_pairs <= st.fallback_pairs
multi = e.compute_multi(cases[-3:])
for b, m in zip(cases[-3:], multi):
    assert np.array_equal(m, e.compute(b))
# qualities outside the regime of the range-extension argument: the guard must route such reads to the fp64 kernel
b = synth.random_batch(37, 40, 10, unrelated=0.5)
b.ins_gop[::2] = 2; b.ins_gop[1::2] = 120; b.gcp[::3] = 45; b.read_quals[::5] = 93
out, ref = e.compute(b), checker(b)
ok = np.isfinite(ref)
assert np.array_equal(np.isfinite(out), ok) and ok.any()
assert float(np.max(np.abs(out[ok] - ref[ok]) / np.abs(ref[ok]))) <= 1e-5
st = e.stats()
assert st.fallback_pairs > 0 and st.fp64_pairs == st.fallback_pairs
print("R2 OK", tot_fb, tot_64)
assert tot_fb > 1000 and tot_64 < 0.5 * tot_fb
"""


def test_range_extended_rerun_matches_the_reference():
    env = dict(os.environ, GKLB_R2="1", PYTHONPATH=str(ROOT))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "R2 OK" in r.stdout
