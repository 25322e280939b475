#!/bin/bash
mkdir -p gpurun_out
for mode in default fine0 fine1 fine2 nochain default; do
  case $mode in default) E="";; fine0) E="TAMF_FINE=0";; fine1) E="TAMF_FINE=1";; fine2) E="TAMF_FINE=2";; nochain) E="TAMF_CHAIN=0";; esac
  echo "== $mode" >> gpurun_out/diag.log
  env $E timeout 300 python tools/diag_refine.py >> gpurun_out/diag.log 2>&1
done
cat gpurun_out/diag.log
