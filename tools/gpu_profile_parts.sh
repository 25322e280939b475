#!/usr/bin/env bash
# One gpurun call: config-3 bench (R: FK + NN + transformer) and ncu --set full captures of the attention, FK and NN kernels.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_profile_parts.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python tools/bench_refine.py > gpurun_out/bench_refine.json 2> gpurun_out/bench_refine.err
tail -n 3 gpurun_out/bench_refine.err; cat gpurun_out/bench_refine.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 8 -c 2 -o gpurun_out/prof_attn -f \
    python bench.py --steps 1 --warmup 1 --chain-steps 4 --no-cpu-baseline --profile-reps 1 > gpurun_out/ncu_attn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mano_fk|h2o_dist|nn_query|vertex_normals' -s 12 -c 8 \
    -o gpurun_out/prof_refine -f python tools/bench_refine.py --reps 1 --warmup 1 > gpurun_out/ncu_refine.log 2>&1
tail -n 3 gpurun_out/ncu_attn.log gpurun_out/ncu_refine.log
ls -la gpurun_out
