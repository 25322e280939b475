#!/usr/bin/env bash
# Weak-scaling bench lines at N = 2, 4, 8 on one box.  Usage: gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_scale.sh 2 4 8'
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for n in "$@"; do
  timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus "$n" --steps 2 --warmup 3 --no-cpu-baseline > "gpurun_out/bench_n$n.json" 2> "gpurun_out/bench_n$n.err"
  tail -n 1 "gpurun_out/bench_n$n.json" | cut -c1-400
done
