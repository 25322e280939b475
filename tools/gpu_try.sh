#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; rm -f gpurun_out/try.log
timeout 1500 python -m pytest tests/test_variants_gpu.py -m gpu -x -q --durations=8 2>&1 | tail -16 >> gpurun_out/try.log
cat gpurun_out/try.log
