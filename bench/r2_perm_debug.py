import os, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gkl_b200 import native, synth, fixtures
b = synth.config2()
e = native.Engine(0, False)
sl = b.read_slice(0, 300)
out = e.compute(sl).reshape(300, b.n_haps)
perm = np.random.default_rng(0).permutation(b.n_haps)
haps = [bytes(b.hap_bases[b.hap_off[h]:b.hap_off[h + 1]]) for h in perm]
off = np.zeros(b.n_haps + 1, dtype=np.int64)
np.cumsum([len(h) for h in haps], out=off[1:])
pb = fixtures.PairHmmBatch(sl.read_off, sl.read_bases, sl.read_quals, sl.ins_gop, sl.del_gop, sl.gcp, off,
                           np.frombuffer(b"".join(haps), dtype=np.uint8).copy())
o2 = e.compute(pb).reshape(300, b.n_haps)
want = out[:, perm]
d = o2 != want
print("mismatches", d.sum(), "of", d.size, "flagged among mismatches", (want[d] < -64.1).sum(), "max abs diff", np.abs(o2 - want).max())
ii = np.argwhere(d)[:10]
for r, h in ii:
    print(r, h, repr(o2[r, h]), repr(want[r, h]), "hap len", len(haps[h]))
os.environ["GKLB_R2"] = "0"
e2 = native.Engine(0, False)
print("fp64 rerun instead: mismatches", (e2.compute(pb).reshape(300, b.n_haps) != e2.compute(sl).reshape(300, b.n_haps)[:, perm]).sum())
