#!/usr/bin/env bash
# Iteration check: GPU parity tests, GEMM timelines, short bench.  Usage: gpurun -- 'bash tools/gpu_iter.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
for w in 0 1 2 3; do timeout 120 python tools/gemm_trace.py $w 2>&1 | head -9; done > gpurun_out/gemm_trace.txt 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/gemm_trace.txt; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
