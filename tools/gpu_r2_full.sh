#!/usr/bin/env bash
# Round-2 full GPU pass: the whole parity suite, smoke, the bench (sample + refine configs) and the in-situ timeline of the layer kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -n 14 > gpurun_out/r2_pytest_gpu.log
tail -n 4 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2_smoke.log 2>&1; tail -n 2 gpurun_out/r2_smoke.log
timeout 200 python tools/chain_trace_model.py 3 > gpurun_out/r2_layer_model_timeline.txt 2>&1
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 1200 gpurun_out/r2_bench.json; tail -n 3 gpurun_out/r2_bench.err
timeout 400 python bench.py --config refine --steps 10 --warmup 3 > gpurun_out/r02_bench_refine_n1.json 2> gpurun_out/r2_bench_refine.err; tail -c 900 gpurun_out/r02_bench_refine_n1.json
