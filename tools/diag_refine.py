"""Debug: per-iteration CUDA-event times of SegmentRefineModel.forward (B=64, T=160, 8192 points)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for q in (ROOT, os.path.join(ROOT, "oakink2-tamf_b200")):
    sys.path.insert(0, q)
import torch
import tamf_b200
from tamf_b200 import synth
cfg = synth.ARCH["arch_refine"]
m = tamf_b200.SegmentRefineModel("unused", **cfg, use_pc=True,
                                 mano_assets={"right": synth.mano_assets("right"), "left": synth.mano_assets("left")})
m.load_state_dict(synth.r_state_dict(cfg, 0), strict=False)
m = m.eval().cuda()
host = synth.make_batch(64, 160, nobj=1, seed=1, npoints=8192, with_pointcloud=True)
batch = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
ts = []
for i in range(40):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = m(batch); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("ms per forward:", " ".join(f"{t:.1f}" for t in ts))
