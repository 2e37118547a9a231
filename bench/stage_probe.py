import sys, os, time
sys.path.insert(0, '/root/repo')
from gkl_b200 import native, synth
regions = synth.config3(32, seed=3)
e = native.Engine(0, False)
outs = e.compute_multi(regions)
os.environ["GKLB_STAGE_TIMING"] = "1"
outs = e.compute_multi(regions, outs)
e.compute(regions[0]); e.compute(regions[0])
