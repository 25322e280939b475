#!/usr/bin/env python
"""BASELINE.json configs[0] as a parity run: MF-MDM G (config/arch_mdm.yml, random init), ONE synthetic two-hand
sequence (rows rh + lh, same text / objects), T=160, nobj=2, the FULL 1000-step ancestral chain -- CUDA path step by
step with explicit per-step noise against the fp32 oracle on the host cores with the SAME noise (synth.step_noise),
free running on both sides.  Prints one JSON line with the drift along the chain.  Takes about a minute of CPU time.

    python tools/parity_full_chain.py [--steps 1000] [--arch arch_mdm]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oakink2-tamf_b200"))
sys.path.insert(0, ROOT)


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--arch", default="arch_mdm")
    ap.add_argument("--frames", type=int, default=160)
    a = ap.parse_args()
    import tamf_b200
    from oracle import tamf_oracle as orc
    from tamf_b200 import synth
    cfg = synth.ARCH[a.arch]
    B, T = 2, a.frames
    batch = synth.make_batch(B, T, nobj=2, seed=0)
    for k in ("text", "obj_list"):  # a two-hand sequence: both rows share task text and objects
        if k in batch:
            batch[k] = [batch[k][0]] * B
    for k in ("shape", "obj_traj", "obj_embedding"):
        batch[k] = batch[k][:1].repeat(B, *([1] * (batch[k].ndim - 1)))
    batch["hand_side"] = ["rh", "lh"]
    shape = (B, 99, 1, T)
    sd = synth.g_state_dict(cfg, seed=0)
    m = tamf_b200.InterationSegmentMDM(**cfg, text_encoder=synth.text_features)
    m.load_state_dict(sd, strict=False)
    m = m.eval().to("cuda")
    dbatch = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    text = synth.text_features(batch["text"])
    tab = orc.diffusion_tables(1000)
    full = tamf_b200.create_gaussian_diffusion(1000, "cosine")
    full._install(m, "ancestral")
    xT = synth.step_noise(123, 1000, shape)
    g, r = xT.cuda(), xT.clone()
    drift = {}
    t0 = time.perf_counter()
    cpu_s = 0.0
    with torch.no_grad():
        for t in range(a.steps - 1, -1, -1):
            n = synth.step_noise(123, t, shape)
            g = m.p_sample_step(g, t, dbatch, noise=n)["sample"]
            c0 = time.perf_counter()
            x0 = orc.g_forward(sd, cfg, r, torch.full((B,), t, dtype=torch.long), batch, text)
            r = orc.p_sample_update(tab, r, x0, t, n)
            cpu_s += time.perf_counter() - c0
            if t in (a.steps - 1, 900, 750, 500, 250, 100, 10, 0):
                drift[str(t)] = rel_l2(g.cpu(), r)
    line = {"what": f"full {a.steps}-step ancestral chain, {a.arch}, one two-hand sequence (B=2: rh + lh), T={T}, nobj=2, "
                    "CUDA bf16 path vs fp32 oracle, same noise, free running (BASELINE.json configs[0])",
            "rel_l2_final": drift["0"], "max_abs_final": float((g.cpu() - r).abs().max()),
            "rel_l2_along_chain": drift, "oracle_cpu_seconds": cpu_s, "oracle_threads": torch.get_num_threads(),
            "host_cpus": os.cpu_count(), "wall_seconds": time.perf_counter() - t0,
            "oracle_two_hand_sequences_per_s": (a.steps / 1000.0) / cpu_s if cpu_s else None}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
