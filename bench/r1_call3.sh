#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pdhmm.py -x -q 2>&1 | tail -5
timeout 600 python bench/pdhmm_bench.py --reads 10000 --haps 128 --iters 3 --cpu-reads 100 --out gpurun_out/pdhmm_v2j.json 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pdhmm -s 1 -c 1 -f -o gpurun_out/prof_pdhmm_v2j \
  python bench/pdhmm_bench.py --reads 1000 --haps 128 --iters 1 --cpu-reads 8 --out gpurun_out/pdhmm_under_ncu.json 2>&1 | tail -2
