set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
cd $GRAFT_REPO_ROOT
for f in test_nn_gpu test_mano_gpu test_gemm_gpu test_denoiser_gpu; do
  timeout 300 python -m pytest tests/$f.py -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/$f.log
  echo "exit $f: $?" >> gpurun_out/$f.log
done
tail -15 gpurun_out/*.log
