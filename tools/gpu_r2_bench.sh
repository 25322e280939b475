#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 400 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 3000 gpurun_out/r2_bench.json; tail -n 5 gpurun_out/r2_bench.err
timeout 300 python bench.py --sequences 128 --no-cpu-baseline > gpurun_out/r2_bench_seq128.json 2> gpurun_out/r2_bench_seq128.err; tail -c 600 gpurun_out/r2_bench_seq128.json; tail -n 5 gpurun_out/r2_bench_seq128.err
timeout 300 python bench.py --config refine > gpurun_out/r2_bench_refine.json 2> gpurun_out/r2_bench_refine.err; tail -c 1500 gpurun_out/r2_bench_refine.json; tail -n 5 gpurun_out/r2_bench_refine.err
