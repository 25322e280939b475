"""Debug aid: per-CTA timeline (clock64) of the layer kernel of ONE encoder layer inside real denoiser steps
(arch_mdm_l, B=64, T=160: L2 state, PDL overlap and neighbours as in the chain).  Run under gpurun:
   python tools/chain_trace_model.py [layer]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oakink2-tamf_b200"))
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
tr = [torch.zeros(148, 64, dtype=torch.int64, device="cuda")]
with m.cond_scope(batch, B, T, x.device):
    for t in range(999, 990, -1):
        x = m.p_sample_step(x, t, batch, seed=1)["sample"]
    _lib.check(_lib.lib().tamf_debug_chain_trace(_lib.ptr(tr[0]), layer), "trace on")
    for t in range(990, 985, -1):
        for z in tr:
            z.zero_()
        x = m.p_sample_step(x, t, batch, seed=1)["sample"]
    torch.cuda.synchronize()
    _lib.check(_lib.lib().tamf_debug_chain_trace(None, -1), "trace off")
for name, z in zip(("layer kernel: LN1 -> L1 -> LN2 -> INP",), tr):
    t = z.cpu().numpy()
    ends = [int(r[3] - r[0]) for r in t if r[0]]
    print(f"{name} (layer {layer}): CTAs {len(ends)}  per-CTA span cycles: min {min(ends)} max {max(ends)}")
    import numpy as np
    live = t[t[:, 0] != 0]
    t0 = live[:, 56].min()
    done = live[:, 55][live[:, 55] != 0] - t0
    seen = live[:, 54][live[:, 54] != 0] - t0
    if len(done) and len(seen):
        last = live[:, 53][live[:, 53] != 0] - t0
        print(f"   globaltimer ns after the earliest dependency-wait end: pdl-wait end spread {int((live[:, 56] - t0).max())}; "
              f"first-LN hi plane complete (warp 0) min/median/max {int(done.min())}/{int(np.median(done))}/{int(done.max())}; "
              f"(slowest warp of a CTA) min/median/max {int(last.min())}/{int(np.median(last))}/{int(last.max())}; "
              f"first dependent unit cleared by the scout min/median/max {int(seen.min())}/{int(np.median(seen))}/{int(seen.max())}")
    for cta in (0, 1, 2, 18, 20, 21, 72, 146, 147):
        r = t[cta]
        if r[0] == 0:
            continue
        z0 = r[0]
        f = lambda v: "   -  " if v == 0 else f"{(v - z0):6d}"
        print(f"CTA {cta}: setup {f(r[1])} pdl {f(r[2])} end {f(r[3])}  first LN unit, warp 0: pass1 {f(r[57])} stats {f(r[58])} "
              f"hi stored {f(r[59])} hi read {f(r[60])} lo stored {f(r[61])} complete {f(r[62])}")
        for it in range(8):
            if r[4 + 6 * it] == 0 and r[8 + 6 * it] == 0:
                break
            print(f"   unit {it}: prod [{f(r[4+6*it])},{f(r[5+6*it])}]  mma [{f(r[6+6*it])},{f(r[7+6*it])}]  epi [{f(r[8+6*it])},{f(r[9+6*it])}]")
