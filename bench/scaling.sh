#!/bin/bash
# Scaling run on one box: bench.py at N = 1, 2, 4, 8 (as the driver launches it) -> gpurun_out/bench_scaling_n<N>.json
mkdir -p gpurun_out
for n in ${GPUS:-1 2 4 8}; do
  if [ "$n" = 1 ]; then
    python bench.py --steps 5 --warmup 3 > gpurun_out/bench_scaling_n1.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) \
      bench.py --gpus $n --steps 5 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/bench_scaling_n$n.json
  fi
  python - "$n" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_scaling_n{sys.argv[1]}.json"))
print(d["n_gpus"], "GPUs:", round(d["value"]), "GCUPS resident,", round(d["e2e"]["value"]), "e2e, roofline",
      round(d["roofline"]["frac"], 3), "cpu", round(d["cpu_baseline"]["value"], 1), "on", d["cpu_baseline"]["cores"],
      "threads", d["clocks"])
PY
done
