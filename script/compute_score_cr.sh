#!/bin/bash
# Contact-ratio score of refined samples (the reference's script/compute_score/compute_score_cr.py).
# usage: script/compute_score_cr.sh <split> <sample_refine dir, e.g. common/sample_refine/main/sample/test/arch_mdm_l__0399> [device id]
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
export PYTHONPATH="$ROOT/oakink2-tamf_b200:$PYTHONPATH"
python -m tamf_b200.launch.compute_score_cr \
    --data.process_range "?(file:./asset/split/$1.txt)" \
    --data.cache_dict_filepath "common/save_cache_dict/main/cache/$1.pkl" \
    --debug.sample_refine_filepath "$2" \
    --runtime.device_id "${3:-0}"
