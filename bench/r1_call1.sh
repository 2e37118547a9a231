#!/bin/bash
# GPU call: two-column sweep variant (VAR 6) against the product kernel, and a source-level profile of the PDHMM kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python bench/sweep.py --reads 10000 --iters 5 --out gpurun_out/sweep_var6.jsonl \
  --variants "list:f2,16,7,12,5;f2,16,7,12,6;f2,16,7,10,6;f2,16,7,8,6;f2,16,7,10,5" 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_pdhmm.py -x -q 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pdhmm -s 1 -c 1 -f -o gpurun_out/prof_pdhmm_v1 \
  python bench/pdhmm_bench.py --reads 1000 --haps 128 --iters 1 --cpu-reads 8 --out gpurun_out/pdhmm_under_ncu.json 2>&1 | tail -3
ls -la gpurun_out
