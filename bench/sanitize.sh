#!/bin/bash
# compute-sanitizer over the newest kernels on small inputs (memcheck: out-of-bounds / misaligned accesses;
# racecheck: shared-memory hazards).  gpurun --timeout 1500 -- bash bench/sanitize.sh
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  echo "== $tool: smith-waterman"
  timeout 600 $CS --tool $tool --print-limit 5 python -m pytest tests/test_gpu_sw.py -q -x -k "known_answers or long_sequences or edge_cases" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|hazard|Error" | head -8
  echo "== $tool: pdhmm"
  timeout 600 $CS --tool $tool --print-limit 5 python -m pytest tests/test_gpu_pdhmm.py -q -x -k "cross_layout or 199_68_51 or 990_1_2" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|hazard|Error" | head -8
done
echo "== memcheck: pairhmm smoke"
timeout 600 $CS --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "ERROR SUMMARY|smoke|Invalid|Error" | head -5
