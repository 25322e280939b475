#!/usr/bin/env bash
# One gpurun call producing everything profiles/ is built from: GPU parity tests, smoke, both bench arms, the config-3
# (refine) bench, the ncu launch list of the bench command, ncu --set full captures of every hot kernel class, and the
# per-CTA timelines.  Usage: gpurun --timeout 1700 -- 'bash tools/gpu_profile_all.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 15 > gpurun_out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python tools/bench_refine.py > gpurun_out/bench_refine.json 2> gpurun_out/bench_refine.err
# launch list of the bench command with a short chain (same kernels per denoiser evaluation)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --chain-steps 8 --no-cpu-baseline \
    --profile-reps 1 > gpurun_out/ncu_bench.log 2>&1
# one full capture of each kernel class of a denoiser evaluation: prep, embed-a, embed-b, one encoder layer; final GEMM
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|attn_tc_kernel|prep_kernel' \
    -s 88 -c 8 -o gpurun_out/prof_denoiser -f python bench.py --steps 1 --warmup 1 --chain-steps 4 --no-cpu-baseline \
    --profile-reps 1 > gpurun_out/ncu_denoiser.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:'Li128ELi5ELi2' \
    -s 2 -c 1 -o gpurun_out/prof_final -f python bench.py --steps 1 --warmup 1 --chain-steps 4 --no-cpu-baseline \
    --profile-reps 1 > gpurun_out/ncu_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'mano_fk|h2o_pruned|nnp_build|vertex_normals' -s 8 -c 5 \
    -o gpurun_out/prof_refine -f python tools/bench_refine.py --reps 1 --warmup 1 > gpurun_out/ncu_refine.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'h2o_scan|nn_scan' -s 0 -c 1 \
    -o gpurun_out/prof_nn_scan -f python tools/bench_refine.py --reps 1 --warmup 1 > gpurun_out/ncu_nn_scan.log 2>&1
timeout 300 python tools/parity_full_chain.py > gpurun_out/parity_full_chain.json 2> gpurun_out/parity_full_chain.err
for w in 0 1 2 3 4 5; do timeout 120 python tools/gemm_trace.py $w; done > gpurun_out/gemm_timelines.txt 2>&1
timeout 120 python tools/attn_trace.py 64 165 512 > gpurun_out/attn_timeline.txt 2>&1
tail -n 4 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.json gpurun_out/bench_ref.json gpurun_out/bench_refine.json \
    gpurun_out/ncu_denoiser.log gpurun_out/ncu_final.log gpurun_out/ncu_refine.log gpurun_out/ncu_nn_scan.log gpurun_out/parity_full_chain.json
ls -la gpurun_out
