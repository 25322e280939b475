#!/usr/bin/env bash
# attention iteration check: parity test + the bench's per-kernel breakdown on a 100-step chain
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attn_gpu.py -x -q 2>&1 | tail -n 8
timeout 600 python bench.py --steps 2 --warmup 3 --chain-steps 100 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/eval', round(j['roofline']['step']['ms_per_denoiser_eval'],4), j['roofline']['kernels_ms'])"
