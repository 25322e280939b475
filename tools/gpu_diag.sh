#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k layer_kernel 2>&1 | tail -3 >> gpurun_out/diag.log
for fine in 3 0; do
  echo "=== TAMF_FINE=$fine" >> gpurun_out/diag.log
  TAMF_FINE=$fine DIAG_N=6000 timeout 300 python tools/diag_repeat.py >> gpurun_out/diag.log 2>&1
  TAMF_FINE=$fine timeout 300 python bench.py --steps 2 --warmup 1 --chain-steps 200 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('bench ms/step', j['ms_per_step'] / 200, j['clocks']['sm_mhz'], j.get('kernels_in_graph_us'))" >> gpurun_out/diag.log 2>&1
done
cat gpurun_out/diag.log
