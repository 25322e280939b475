#!/bin/bash
mkdir -p gpurun_out
for cfg in "arch_mdm 5 24" "arch_mdm 1 171" "arch_mdm 64 160" "arch_mdm_l 3 40"; do
  set -- $cfg
  for E in "" "TAMF_CHAIN=0"; do
    echo "== $cfg $E" >> gpurun_out/diag.log
    env $E DIAG_ARCH=$1 DIAG_B=$2 DIAG_T=$3 DIAG_N=3000 timeout 300 python tools/diag_repeat.py 2>&1 | tail -3 >> gpurun_out/diag.log
  done
done
cat gpurun_out/diag.log
