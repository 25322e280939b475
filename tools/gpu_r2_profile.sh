#!/usr/bin/env bash
# Round-2 profile pass (one gpurun call, 1 GPU): both bench arms, the refine config, the ncu launch list of the bench
# command, ncu --set full captures of one denoiser evaluation's kernel classes, the in-situ layer-kernel timeline.
# Usage: gpurun --timeout 2400 -- 'bash tools/gpu_r2_profile.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r02_gpu.txt
nproc >> gpurun_out/r02_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/r02_gpu.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err
timeout 600 python bench.py --config refine --steps 10 --warmup 3 > gpurun_out/r02_bench_refine_n1.json 2> gpurun_out/r02_bench_refine_n1.err
TAMF_CHAIN=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_round1_form.json 2> /dev/null
timeout 200 python tools/chain_trace_model.py 3 > gpurun_out/r02_layer_kernel_timeline.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv \
    --log-file gpurun_out/r02_launches_bench_chain8.csv python bench.py --steps 1 --warmup 1 --chain-steps 8 --no-cpu-baseline \
    --profile-reps 1 > gpurun_out/r02_ncu_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'gemm_tc_kernel|attn_tc_kernel|prep_kernel|layer_chain_kernel' \
    -s 42 -c 9 -o gpurun_out/r02_prof_denoiser -f python bench.py --steps 1 --warmup 1 --chain-steps 4 --no-cpu-baseline \
    --profile-reps 1 > gpurun_out/r02_ncu_denoiser.log 2>&1
bash tools/ncu_summary.sh gpurun_out/r02_prof_denoiser.ncu-rep > gpurun_out/r02_ncu_denoiser_full_summary.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
# keep the merge under the 64 MiB limit: the report itself only if small
find gpurun_out -name '*.ncu-rep' -size +40M -delete
tail -c 600 gpurun_out/r02_bench_n1.json; tail -c 400 gpurun_out/r02_bench_ref_n1.json; tail -c 400 gpurun_out/r02_bench_refine_n1.json
tail -n 3 gpurun_out/r02_ncu_denoiser.log; head -c 1500 gpurun_out/r02_ncu_denoiser_full_summary.csv
