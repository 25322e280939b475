#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k layer 2>&1 | tail -n 5 > gpurun_out/r2_chain_unit.log
if ! grep -q "passed" gpurun_out/r2_chain_unit.log || grep -q "failed" gpurun_out/r2_chain_unit.log; then echo "layer unit test failed: stop"; tail gpurun_out/r2_chain_unit.log; exit 1; fi
timeout 200 python tools/chain_trace_model.py 3 > gpurun_out/r2_chain_model_timelines.txt 2>&1
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --profile-reps 3"
run() { tag=$1; shift; env "$@" timeout 300 $B > gpurun_out/r2_b_$tag.json 2> gpurun_out/r2_b_$tag.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_b_$tag.json")); print("$tag", round(d["value"],2), "seq/s", round(d["roofline"]["step"]["ms_per_denoiser_eval"],4), "ms", d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"), d["roofline"]["kernels_ms"])
except Exception as e: print("$tag failed", e)
PY
}
run default X=1
run pdl0 TAMF_PDL=0

run nochain TAMF_CHAIN=0
