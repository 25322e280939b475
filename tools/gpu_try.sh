#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; rm -f gpurun_out/try.log
timeout 900 python -m pytest tests/test_launch_gpu.py -m gpu -x -q 2>&1 | tail -15 >> gpurun_out/try.log
cat gpurun_out/try.log
