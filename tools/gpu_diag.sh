#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k layer_kernel 2>&1 | tail -3 >> gpurun_out/diag.log
timeout 900 python -m pytest tests/test_denoiser_gpu.py tests/test_refine_gpu.py tests/test_launch_gpu.py -m gpu -x -q 2>&1 | tail -5 >> gpurun_out/diag.log
DIAG_N=4000 timeout 300 python tools/diag_repeat.py 2>&1 | tail -2 >> gpurun_out/diag.log
for i in 1 2; do
timeout 300 python bench.py --steps 2 --warmup 1 --chain-steps 200 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('bench ms/step', j['ms_per_step'] / 200, j['clocks']['sm_mhz'], j['roofline'].get('kernels_in_graph_us'))" >> gpurun_out/diag.log 2>&1
done
cat gpurun_out/diag.log
