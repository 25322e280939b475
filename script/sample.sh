#!/bin/bash
# Sample a split with MF-MDM G on B200s (the reference's script/sample.sh, without the interactive prompt).
# usage: script/sample.sh <split> <model weights .pt> <model name> [device ids, default 0,1,2,3,4,5,6,7]
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
export PYTHONPATH="$ROOT/oakink2-tamf_b200:$PYTHONPATH"
python -m tamf_b200.launch.sample \
    --cfg "$ROOT/config/obj_embedding.yml" \
    --data.process_range "?(file:./asset/split/$1.txt)" \
    --data.cache_dict_filepath "common/save_cache_dict/main/cache/$1.pkl" \
    --cfg "$ROOT/config/arch_mdm_l.yml" \
    --debug.model_weight_filepath "$2" \
    --debug.sample_save_offset "$1/$3" \
    --runtime.device_id "${4:-0,1,2,3,4,5,6,7}" \
    --commit
