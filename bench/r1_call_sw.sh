#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sw.py tests/test_jni.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench/sw_bench.py --out gpurun_out/sw_bench_v3.json 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_smith -s 1 -c 1 -f -o gpurun_out/prof_sw_v3 \
  python bench/sw_bench.py --pairs 4000 --reps 1 --out gpurun_out/sw_under_ncu.json 2>&1 | tail -2
