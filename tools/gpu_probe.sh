#!/usr/bin/env bash
# One gpurun call: TMA ingest probe, per-CTA GEMM timelines, ncu full capture of the attention kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
(cd tools/probes && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_bw tma_bw.cu -lcuda && timeout 120 /tmp/tma_bw) > gpurun_out/tma_bw.txt 2>&1
for w in 0 1 2 3; do timeout 120 python tools/gemm_trace.py $w; done > gpurun_out/gemm_trace.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 8 -c 2 \
    -o gpurun_out/prof_attn -f python bench.py --steps 1 --warmup 1 --chain-steps 3 --no-cpu-baseline --profile-reps 1 \
    > gpurun_out/prof_attn.log 2>&1
tail -40 gpurun_out/gemm_trace.txt
