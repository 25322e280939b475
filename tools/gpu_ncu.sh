#!/usr/bin/env bash
# ncu full capture of selected kernels inside a short bench run.  Usage: bash tools/gpu_ncu.sh <regex> <skip> <count> <outname>
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$1" -s "$2" -c "$3" \
    -o "gpurun_out/$4" -f python bench.py --steps 1 --warmup 1 --chain-steps 3 --no-cpu-baseline --profile-reps 1 \
    > "gpurun_out/$4.log" 2>&1
tail -3 "gpurun_out/$4.log"
