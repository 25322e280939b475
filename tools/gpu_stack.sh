#!/bin/bash
# Stack form (TAMF_CHAIN=2, experimental): parity + determinism + short benches over the SM split; the default form beside it.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; rm -f gpurun_out/stack.log
TAMF_CHAIN=2 timeout 900 python -m pytest tests/test_denoiser_gpu.py tests/test_refine_gpu.py -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/stack.log
for E in "TAMF_CHAIN=2 TAMF_STACK_ATT=20" "TAMF_CHAIN=2 TAMF_STACK_ATT=24" "TAMF_CHAIN=2 TAMF_STACK_ATT=32" "TAMF_CHAIN=1"; do
  env $E timeout 300 python bench.py --steps 2 --warmup 1 --chain-steps 200 --no-cpu-baseline 2>gpurun_out/stack_bench.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('$E', 'ms/step', round(j['ms_per_step'] / 200, 4), j['clocks']['sm_mhz'], j['roofline'].get('in_graph_step_us'))" >> gpurun_out/stack.log 2>&1
  grep -i "error\|Traceback" gpurun_out/stack_bench.err | tail -2 >> gpurun_out/stack.log
done
cat gpurun_out/stack.log
