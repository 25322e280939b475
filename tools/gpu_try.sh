#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; rm -f gpurun_out/try.log
for E in "TAMF_CHAIN_DBG=0" "TAMF_CHAIN_DBG=32" "TAMF_CHAIN_DBG=0" "TAMF_CHAIN_DBG=32" "TAMF_CHAIN_DBG=1" "TAMF_CHAIN_DBG=33"; do
env $E timeout 300 python bench.py --steps 2 --warmup 1 --chain-steps 300 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j = json.loads(l); print('$E', 'ms/step', round(j['ms_per_step'] / 300, 4), j['clocks']['sm_mhz'], j['clocks']['power_w_max'], j['roofline'].get('in_graph_step_us'))" >> gpurun_out/try.log 2>&1
done
cat gpurun_out/try.log
