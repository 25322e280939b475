#!/usr/bin/env bash
# Summarise an ncu --set full report into the CSV kept under profiles/ (one row per captured launch).
# Usage: tools/ncu_summary.sh gpurun_out/prof_denoiser.ncu-rep > profiles/rNN_ncu_denoiser_full_summary.csv
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,sm__cycles_elapsed.max,smsp__cycles_active.avg,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"
ncu -i "$1" --page raw --csv --metrics $M 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); h = rows[0]
keep = [i for i, c in enumerate(h) if c in ('Kernel Name', 'Grid Size', 'Block Size') or '.' in c or c.startswith('launch__')]
w = csv.writer(sys.stdout)
for r in rows: w.writerow([r[i] for i in keep])
"
