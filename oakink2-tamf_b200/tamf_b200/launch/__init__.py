"""Launchers: `python -m tamf_b200.launch.sample` and `python -m tamf_b200.launch.sample_refine`, the entry points behind
the reference's script/sample.sh and script/sample_refine.sh (src/oakink2_tamf/launch/sample.py, sample_refine.py), with
the same flags, YAML configs and output layout."""
