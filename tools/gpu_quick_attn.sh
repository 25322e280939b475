#!/usr/bin/env bash
# Attention iteration check: GPU parity tests + short bench, then the per-CTA attention timeline with every exp2 on the
# MUFU unit (default) and with the polynomial share on the FMA pipe (TAMF_ATTN_DBG=2).
# Usage: gpurun -- 'bash tools/gpu_quick_attn.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
bash tools/gpu_quick.sh
for dbg in 0 2; do
  echo "== TAMF_ATTN_DBG=$dbg"
  TAMF_ATTN_DBG=$dbg timeout 120 python tools/attn_trace.py 64 165 512 2>&1 | tee gpurun_out/attn_timeline_dbg$dbg.txt | grep -E "CTA   0|per-CTA|event"
done
TAMF_ATTN_DBG=2 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TAMF_ATTN_DBG=2 seq/s', round(j['value'],2), j['roofline']['kernels_ms']['attention'], j['clocks'])"
