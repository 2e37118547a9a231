"""The 32 config-3 regions as one multi-region job, for ncu captures of the multi-class kernels (GPU box):
    ncu --set full --clock-control none --import-source on -k regex:k_h2_mega -s 1 -c 1 -o gpurun_out/mega python bench/profile_c3.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gkl_b200 import native, synth

regions = synth.config3(32, seed=3)
e = native.Engine(0, False)
for _ in range(3):
    e.compute_multi(regions)
print(e.sweep_kernel(), e.stats().kernel_ms, "ms of kernels")
e.close()
