#!/usr/bin/env bash
# Round-2 multi-GPU pass on one box: for every N given: weak scaling (64 sequences per GPU), the reference arm under
# torchrun (rank 0 only; must end within its time budget), strong scaling (--sequences S fixed over N ranks), and the
# sample launcher over N devices.  Usage: gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_r2_scale.sh 2'
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
SEQ=${SEQ:-512}
for n in "$@"; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n))"
  timeout 400 $TR bench.py --gpus "$n" --steps 2 --warmup 3 --no-cpu-baseline > "gpurun_out/r02_bench_n$n.json" 2> "gpurun_out/r02_bench_n$n.err"
  tail -n 1 "gpurun_out/r02_bench_n$n.json" | cut -c1-300
  timeout 400 $TR bench.py --gpus "$n" --impl reference --steps 1 --warmup 1 --ref-budget-s 60 > "gpurun_out/r02_bench_ref_n$n.json" 2> "gpurun_out/r02_bench_ref_n$n.err"
  echo "reference arm rc=$?"; tail -n 1 "gpurun_out/r02_bench_ref_n$n.json" | cut -c1-300
  timeout 900 $TR bench.py --gpus "$n" --sequences "$SEQ" --steps 1 --warmup 1 --no-cpu-baseline > "gpurun_out/r02_bench_strong${SEQ}_n$n.json" 2> "gpurun_out/r02_bench_strong${SEQ}_n$n.err"
  tail -n 1 "gpurun_out/r02_bench_strong${SEQ}_n$n.json" | cut -c1-400
  rm -rf /tmp/tamf_launch && mkdir -p /tmp/tamf_launch && (cd /tmp/tamf_launch && PYTHONPATH="$OLDPWD/oakink2-tamf_b200" timeout 300 python -m tamf_b200.launch.sample \
      --cfg "$OLDPWD/config/arch_mdm_l.yml" --data.source synthetic:$((8 * n)):40:2 --debug.random_init_seed 0 --runtime.text_encoder synthetic \
      --runtime.device_id "$(seq -s, 0 $((n - 1)))" --runtime.seed 3 --debug.sample_save_offset test/x --commit > "$OLDPWD/gpurun_out/r02_launch_n$n.log" 2>&1; \
      echo "launcher rc=$? files=$(ls common/sample/main/sample/test/x | wc -l)")
done
