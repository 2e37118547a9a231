timeout 300 python -m pytest tests/test_gpu_pdhmm.py -m gpu -q --timeout 300 2>&1 | tail -2
for v in 1 2 x; do
echo "variant $v"
GKLB_PDHMM_V3=$v GKLB_PDHMM_KERNEL=$([ $v = x ] && echo 2 || echo 3) timeout 300 python bench/pdhmm_bench.py --reads 6000 --read-len 150 --cpu-reads 30 --iters 3 --out gpurun_out/pd150_$v.json 2>&1 | tail -1 | python -c "
import sys, json; r = json.loads(sys.stdin.read()); print(r['gpu_kernel_gcups'], r['gpu_kernel_ms'], r['max_abs_diff_vs_serial_restatement'])"
done
