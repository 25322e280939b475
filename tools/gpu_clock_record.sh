#!/bin/bash
# Clock / power record of the timed chain (100 ms samples of nvidia-smi while bench.py runs), default form and round-1 form.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
Q="timestamp,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown"
for form in default round1; do
  out=gpurun_out/r02_clock_record_$form.csv
  nvidia-smi --query-gpu=$Q --format=csv -lms 100 -i 0 > $out &
  SMI=$!
  sleep 1
  if [ $form = round1 ]; then E="TAMF_CHAIN=0"; else E="TAMF_CHAIN=1"; fi
  env $E timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_clock_bench_$form.json 2>/dev/null
  sleep 1
  kill $SMI
  python - <<PY
import csv, json, statistics as st
rows = list(csv.reader(open("$out")))[1:]
busy = [r for r in rows if len(r) > 4 and float(r[3].split()[0]) > 600]
clk = [float(r[1].split()[0]) for r in busy]; pw = [float(r[3].split()[0]) for r in busy]
cap = sum(1 for r in busy if r[6].strip() == "Active")
j = [json.loads(l) for l in open("gpurun_out/r02_clock_bench_$form.json") if l.startswith("{")][0]
print("$form: samples under load %d  SM clock median %.0f min %.0f max %.0f MHz (max clock %s)  power median %.0f W (limit %s)  sw_power_cap active in %d samples  temp max %s C  -> %.2f seq/s, %.1f ms per chain" % (
    len(busy), st.median(clk), min(clk), max(clk), rows[0][2].strip(), st.median(pw), rows[0][4].strip(), cap, max(r[5].strip() for r in busy), j["value"], j["ms_per_step"]))
PY
done
