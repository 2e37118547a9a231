#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pdhmm.py -x -q 2>&1 | tail -5
timeout 600 python bench/pdhmm_bench.py --reads 10000 --haps 128 --iters 3 --cpu-reads 100 --out gpurun_out/pdhmm_c5_final.json 2>&1 | tail -1 | cut -c1-300
timeout 600 python bench/pdhmm_bench.py --reads 10000 --haps 128 --read-len 150 --iters 3 --cpu-reads 50 --out gpurun_out/pdhmm_150_v2.json 2>&1 | tail -1 | cut -c1-300
GKLB_PDHMM_KERNEL=1 timeout 600 python bench/pdhmm_bench.py --reads 10000 --haps 128 --read-len 150 --iters 3 --cpu-reads 8 --out gpurun_out/pdhmm_150_v1.json 2>&1 | tail -1 | cut -c1-300
