#!/usr/bin/env python
"""Strided / DDIM sampling at the configs[1] size (arch_mdm_l, B=64, T=160): sequences/s of `ddim_sample_loop` for a
few step counts, next to the 1000-step ancestral chain.  NOT the headline metric (the reference launchers sample with
the full ancestral chain); it shows what the samplers the reference file offers cost on this path.

    python tools/bench_ddim.py [--specs ddim50,ddim100,ddim250] [--eta 0.0]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oakink2-tamf_b200"))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--specs", default="ddim50,ddim100,ddim250")
    ap.add_argument("--eta", type=float, default=0.0)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import tamf_b200
    from tamf_b200 import synth
    dev = torch.device("cuda:0")
    cfg = synth.ARCH["arch_mdm_l"]
    B, T = a.batch, 160
    m = tamf_b200.InterationSegmentMDM(**cfg, text_encoder=synth.text_features)
    m.load_state_dict(synth.g_state_dict(cfg, seed=0), strict=False)
    m = m.eval().to(dev)
    batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.make_batch(B, T, nobj=2, seed=0).items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    for spec in ["1000"] + a.specs.split(","):
        d = tamf_b200.create_gaussian_diffusion(1000, "cosine", timestep_respacing="" if spec == "1000" else spec)
        fn = (lambda: d.p_sample_loop(m, (B, 99, 1, T), clip_denoised=False, model_kwargs={"batch": batch}, seed=1)) \
            if spec == "1000" else \
            (lambda: d.ddim_sample_loop(m, (B, 99, 1, T), clip_denoised=False, model_kwargs={"batch": batch}, eta=a.eta,
                                        seed=1))
        x = fn()
        assert torch.isfinite(x).all()
        ms = []
        for _ in range(a.reps):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        best = sum(ms) / len(ms)
        out[spec] = {"steps": d.num_timesteps, "ms_per_batch": best, "sequences_per_s": B / (best * 1e-3),
                     "ms_per_step": best / d.num_timesteps}
    print(json.dumps({"what": f"arch_mdm_l, B={B}, T={T}: sampler step counts (1000 = ancestral p_sample_loop; others "
                              f"ddim_sample_loop eta={a.eta}); includes x_T draw and conditioning", "results": out}))


if __name__ == "__main__":
    main()
