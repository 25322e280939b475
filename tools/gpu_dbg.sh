#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for d in 0 4 8; do echo "=== which=3 DBG=$d"; TAMF_GEMM_DBG=$d timeout 120 python tools/gemm_trace.py 3 2>/dev/null | head -3; done
for w in 0 1 2; do timeout 120 python tools/gemm_trace.py $w 2>/dev/null | head -8; done
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['roofline']['kernels_ms'])"
