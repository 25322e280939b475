#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 200 python tools/chain_trace_model.py 3 > gpurun_out/r2_chain_model_timelines.txt 2>&1
grep "CTA" gpurun_out/r2_chain_model_timelines.txt | head -30
