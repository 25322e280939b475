#!/bin/bash
# compute-sanitizer over the layer kernel (one small launch) and a small denoiser step: memcheck, then racecheck.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; rm -f gpurun_out/sanitize.log
for tool in memcheck racecheck; do
  echo "== $tool: layer kernel unit test (M=300, d=256)" >> gpurun_out/sanitize.log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k "layer_kernel and 300-256" -p no:cacheprovider 2>&1 | tail -12 >> gpurun_out/sanitize.log
done
echo "== memcheck: small denoiser forward (arch_mdm edge shapes)" >> gpurun_out/sanitize.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_denoiser_gpu.py -m gpu -x -q -k "edge_shapes" -p no:cacheprovider 2>&1 | tail -12 >> gpurun_out/sanitize.log
cat gpurun_out/sanitize.log
