#!/usr/bin/env bash
# Round-2 full GPU pass: the whole parity suite, smoke, the bench (both arms) and the in-situ timeline of the layer kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -n 14 > gpurun_out/r2_pytest_gpu.log
tail -n 4 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2_smoke.log 2>&1; tail -n 2 gpurun_out/r2_smoke.log
timeout 200 python tools/chain_trace_model.py 3 > gpurun_out/r2_layer_model_timeline.txt 2>&1
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 1800 gpurun_out/r2_bench.json; tail -n 3 gpurun_out/r2_bench.err
