#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; rm -f gpurun_out/try.log
for E in "TAMF_CHAIN_OOO=0" "TAMF_CHAIN_OOO=8 TAMF_CHAIN_DBG=64" "TAMF_CHAIN_OOO=8 TAMF_CHAIN_DBG=192" "TAMF_CHAIN_OOO=2 TAMF_CHAIN_DBG=192" "TAMF_CHAIN_OOO=1 TAMF_CHAIN_DBG=192"; do
env $E timeout 300 python bench.py --steps 2 --warmup 1 --chain-steps 200 --no-cpu-baseline 2>gpurun_out/try_bench.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('$E', 'ms/step', round(j['ms_per_step'] / 200, 4), j['clocks']['sm_mhz'], j['roofline'].get('in_graph_step_us'))" >> gpurun_out/try.log 2>&1
grep -i "error" gpurun_out/try_bench.err | tail -1 >> gpurun_out/try.log
done
TAMF_CHAIN_OOO=8 TAMF_CHAIN_DBG=208 timeout 300 python tools/stack_stalls.py 3 >> gpurun_out/try.log 2>&1
cat gpurun_out/try.log
