#!/usr/bin/env bash
# Iteration check: GPU parity tests + short bench with the per-kernel breakdown.  Usage: gpurun -- 'bash tools/gpu_quick.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 15 > gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -n 6 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("seq/s", round(j["value"], 2), "e2e", round(j["e2e"]["value"], 2), "ms/eval", round(j["roofline"]["step"]["ms_per_denoiser_eval"], 4),
      "frac", round(j["roofline"]["step"]["frac"], 4))
print(j["roofline"]["kernels_ms"])
print(j["clocks"])
PY
