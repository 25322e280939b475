"""Debug aid: where the layer kernel's warps wait (cycles accumulated per CTA over ONE launch inside real denoiser steps,
TAMF_CHAIN_DBG=16): producer on dependencies / on a free ring slot, MMA warp on loaded stages / on a free accumulator,
epilogue warp 0 on a finished accumulator / on the partner half's LayerNorm statistics.
   TAMF_CHAIN=2 TAMF_CHAIN_DBG=16 python tools/stack_stalls.py      (stack form: the one launch covers all layers)
   TAMF_CHAIN_DBG=16 python tools/stack_stalls.py 3                 (per-layer form: layer 3)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oakink2-tamf_b200"))
import numpy as np
import torch
import tamf_b200
from tamf_b200 import _lib, synth

layer = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = synth.ARCH["arch_mdm_l"]
m = tamf_b200.InterationSegmentMDM(**cfg, text_encoder=synth.text_features)
m.load_state_dict(synth.g_state_dict(cfg, seed=0), strict=False)
m = m.eval().cuda()
B, T = 64, 160
batch = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in synth.make_batch(B, T, nobj=2, seed=0).items()}
x = torch.randn(B, 99, 1, T, device="cuda")
tamf_b200.create_gaussian_diffusion(1000, "cosine")._install(m, "ancestral")
G = int(sys.argv[2]) if len(sys.argv) > 2 else 148  # CTAs of the launch (stack form: 2 x its pairs, e.g. 120)
tr = torch.zeros(G * 64 + G * 8, dtype=torch.int64, device="cuda")
with m.cond_scope(batch, B, T, x.device):
    for t in range(999, 990, -1):
        x = m.p_sample_step(x, t, batch, seed=1)["sample"]
    _lib.check(_lib.lib().tamf_debug_chain_trace(_lib.ptr(tr), layer), "trace on")
    for t in range(990, 987, -1):
        tr.zero_()
        x = m.p_sample_step(x, t, batch, seed=1)["sample"]
    torch.cuda.synchronize()
    _lib.check(_lib.lib().tamf_debug_chain_trace(None, -1), "trace off")
t = tr.cpu().numpy()
stamps, st = t[:G * 64].reshape(G, 64), t[G * 64:].reshape(G, 8)
live = stamps[:, 0] != 0
n = int(live.sum())
total = (st[live, 6] - stamps[live, 0]).astype(np.float64)
names = ["producer: dependencies", "producer: free ring slot", "MMA warp: loaded stage", "MMA warp: free accumulator",
         "epilogue warp 0: finished accumulator", "epilogue warp 0: partner statistics"]
print(f"CTAs {n}  kernel cycles per CTA: mean {total.mean():.0f} min {total.min():.0f} max {total.max():.0f}; units per CTA "
      f"mean {st[live, 7].mean():.1f}")
for k, nm in enumerate(names):
    v = st[live, k].astype(np.float64)
    print(f"  {nm:40s} mean {v.mean():9.0f} cycles = {100 * v.mean() / total.mean():5.1f} % of the kernel  (max {v.max():.0f})")
