"""One config-2-shaped run of the engine for ncu captures (GPU box):
    ncu --set full --clock-control none --import-source on -k regex:k_h2 -c 1 -o gpurun_out/prof python bench/profile_c2.py [reads [read_len [haps]]]
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gkl_b200 import native, synth

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
read_len = int(sys.argv[2]) if len(sys.argv) > 2 else 101
haps = int(sys.argv[3]) if len(sys.argv) > 3 else 128
b = synth.config2(reads, haps, read_len)
e = native.Engine(0, False)
e.stage(b)
for _ in range(2):
    e.run()
e.synchronize()
print(e.sweep_kernel(), e.time_runs(3), "ms per run")
e.close()
