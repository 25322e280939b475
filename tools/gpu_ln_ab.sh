#!/usr/bin/env bash
# A/B of the LayerNorm tile height (TAMF_LN_RQ rows per TMEM lane quarter): GPU parity tests with the default, then the
# short bench at 32 (dense 128-row tiles), 24 and the automatic value.  Usage: gpurun -- 'bash tools/gpu_ln_ab.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 15 > gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
for rq in 32 24 0; do
  TAMF_LN_RQ=$rq timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rq$rq.json 2> gpurun_out/bench_rq$rq.err
  tail -n 3 gpurun_out/bench_rq$rq.err
  python - "$rq" <<'PY'
import json, sys
rq = sys.argv[1]
j = json.loads(open(f"gpurun_out/bench_rq{rq}.json").read().strip().splitlines()[-1])
print("RQ", rq, "seq/s", round(j["value"], 2), "e2e", round(j["e2e"]["value"], 2), "ms/eval", round(j["roofline"]["step"]["ms_per_denoiser_eval"], 4))
print("   ", j["roofline"]["kernels_ms"]); print("   ", j["clocks"])
PY
done
for w in 2 3; do timeout 120 python tools/gemm_trace.py $w; done > gpurun_out/gemm_timelines_ln.txt 2>&1
head -n 12 gpurun_out/gemm_timelines_ln.txt
