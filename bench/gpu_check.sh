#!/bin/bash
# What the driver runs at round end, in one gpurun call: GPU tests, smoke(), both bench arms.
#   gpurun --timeout 2400 -- bash bench/gpu_check.sh
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-400
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_n1.json | cut -c1-1500
