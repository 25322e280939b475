#!/bin/bash
# Refine generated samples with MF-MDM R (the reference's script/sample_refine.sh, without the interactive prompt).
# usage: script/sample_refine.sh <split> <refine weights .pt> <model name> [generated sample dir] [device id]
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
export PYTHONPATH="$ROOT/oakink2-tamf_b200:$PYTHONPATH"
python -m tamf_b200.launch.sample_refine \
    --data.process_range "?(file:./asset/split/$1.txt)" \
    --data.cache_dict_filepath "common/save_cache_dict/main/cache/$1.pkl" \
    --data.sample_dir "${4:-common/sample/main/sample/test/arch_mdm_l__0399}" \
    --debug.model_weight_filepath "$2" \
    --debug.sample_save_offset "$1/$3" \
    --runtime.device_id "${5:-0}" \
    --commit
