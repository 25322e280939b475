#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k "layer_kernel and 300-256" -p no:cacheprovider > gpurun_out/racecheck_full.log 2>&1
grep -n "layer_chain.cuh\|Race reported\|hazard" gpurun_out/racecheck_full.log | head -80
