#!/usr/bin/env bash
# Round-2 GPU pass for the chain kernels: numerics of the kernel alone, the whole parity suite, per-CTA timelines, and the
# bench with / without the chain form.  Usage: gpurun --timeout 1500 -- 'bash tools/gpu_r2_chain.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k chain 2>&1 | tail -n 15 > gpurun_out/r2_chain_unit.log
tail -n 5 gpurun_out/r2_chain_unit.log
if ! grep -q "passed" gpurun_out/r2_chain_unit.log || grep -q "failed" gpurun_out/r2_chain_unit.log; then echo "chain unit test failed: stop"; exit 1; fi
for w in 0 1 2; do timeout 120 python tools/chain_trace.py $w; done > gpurun_out/r2_chain_timelines.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 15 > gpurun_out/r2_pytest_b.log
tail -n 5 gpurun_out/r2_pytest_b.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_chain.json 2> gpurun_out/r2_bench_chain.err
TAMF_CHAIN=0 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_nochain.json 2> gpurun_out/r2_bench_nochain.err
tail -c 1500 gpurun_out/r2_bench_chain.json; echo; tail -c 600 gpurun_out/r2_bench_nochain.json; tail -n 3 gpurun_out/r2_bench_chain.err
