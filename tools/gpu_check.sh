#!/usr/bin/env bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list + one full capture of the top kernel.
# Usage (from the authoring container): gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [quick]'
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
if [[ "${1:-}" != "quick" ]]; then
  # launch list of the bench command with a short chain (same kernels per denoiser evaluation)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --chain-steps 8 --no-cpu-baseline \
      --profile-reps 1 > gpurun_out/ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 60 -c 6 \
      -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 1 --chain-steps 4 --no-cpu-baseline \
      --profile-reps 1 > gpurun_out/ncu_full.log 2>&1
fi
tail -n 5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.json gpurun_out/bench.err gpurun_out/bench_ref.json
