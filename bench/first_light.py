"""First-light check on the GPU box: golden vectors, config 1, ragged/adversarial batches, long reads."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import oracle
from gkl_b200 import fixtures, native, synth
from gkl_b200.batch import PairHmmBatch


def check(name, e, b, use_double=False, tol=1e-5):
    out = e.compute(b)
    ref = oracle.ref_pairhmm(b, use_double, threads=oracle.host_threads())[0] if oracle.ref_available() else \
        oracle.port_pairhmm(b, use_double, threads=oracle.host_threads())[0]
    err = np.abs(out - ref) / np.maximum(np.abs(ref), 1e-30)
    st = e.stats()
    print(f"{name:28s} pairs={st.pairs:8d} fb={st.fallback_pairs:7d} classes={st.n_classes} launches={st.kernel_launches} "
          f"max_rel={np.nanmax(err):.3e} nan={int(np.isnan(out).sum())} {'OK' if np.nanmax(err) <= tol and not np.isnan(out).any() else 'FAIL'}",
          flush=True)
    return out, ref


print(native.lib().gklb_version().decode(), "devices:", native.device_count())
e = native.Engine(0, False)
ed = native.Engine(0, True)
bs, exp = fixtures.load_pairhmm_testdata(Path(__file__).resolve().parents[1] / "tests/golden/pairhmm-testdata.txt")
got = np.array([e.compute(b)[0] for b in bs])
gotd = np.array([ed.compute(b)[0] for b in bs])
print("golden fp32 max abs", np.abs(got - exp).max(), " fp64 max abs", np.abs(gotd - exp).max())
b, kat = fixtures.simple_test_batch()
print("simpleTest", e.compute(b)[0], kat)
check("config1", e, synth.config1())
check("config2 500x128", e, synth.config2(500, 128))
check("config2 500x128 double", ed, synth.config2(500, 128), True)
check("random ragged", e, synth.random_batch(11, 300, 40))
check("random lowq+unrelated", e, synth.random_batch(12, 300, 40, low_quality=0.1, unrelated=0.3))
check("random double", ed, synth.random_batch(13, 100, 30, low_quality=0.1, unrelated=0.3), True)
check("long reads (multi-pass)", e, synth.random_batch(14, 40, 12, read_len=(200, 700), hap_len=(300, 900)))
check("long reads double", ed, synth.random_batch(15, 20, 8, read_len=(200, 700), hap_len=(300, 900)), True)
check("tiny", e, synth.random_batch(16, 7, 3, read_len=(1, 12), hap_len=(1, 12)))
check("many haps (tiles)", e, synth.random_batch(17, 30, 700, read_len=(50, 120), hap_len=(200, 450)))
for i, reg in enumerate(synth.config3(3)):
    check(f"config3 region {i}", e, reg)
