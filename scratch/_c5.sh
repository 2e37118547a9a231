timeout 300 python -m pytest tests/test_gpu_pdhmm.py -m gpu -q --timeout 300 2>&1 | tail -2
timeout 300 python -c "
import json, importlib.util
spec = importlib.util.spec_from_file_location('cfgs', 'bench/configs.py'); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
m.PEAKS.update(fp32=lambda: (70.9, 'x'), fp64=lambda: (33.8, 'x'))
r = m.config5(); print(r['value'], r['e2e']['value'], r['parity']['max_abs_err'], r['gpu_launches'], r['roofline']['kernel'])
" 2>&1 | tail -1
timeout 300 python bench/pdhmm_bench.py --reads 6000 --read-len 150 --cpu-reads 30 --iters 3 --out gpurun_out/pd150.json 2>&1 | tail -1 | python -c "
import sys, json; r = json.loads(sys.stdin.read()); print(r['gpu_kernel_gcups'], r['gpu_kernel_ms'], r['max_abs_diff_vs_serial_restatement'])"
