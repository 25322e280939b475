#!/bin/bash
mkdir -p gpurun_out
ls oracle/_ref | head -5 >> gpurun_out/diag.log
timeout 500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err
tail -c 900 gpurun_out/r02_bench_ref_n1.json >> gpurun_out/diag.log; tail -3 gpurun_out/r02_bench_ref_n1.err >> gpurun_out/diag.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python -c "
import json
for l in open('gpurun_out/r02_bench_n1.json'):
    if l.startswith('{'):
        j=json.loads(l); print('ours', j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks']['sm_mhz'], j['cpu_baseline'])" >> gpurun_out/diag.log
cat gpurun_out/diag.log
