"""Debug aid: per-CTA timeline (clock64) of one launch of the tcgen05 attention kernel.  Run under gpurun:
   python tools/attn_trace.py [B] [S] [d]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oakink2-tamf_b200"))
import numpy as np
import torch
from tamf_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 165
d = int(sys.argv[3]) if len(sys.argv) > 3 else 512
H = 4
L = _lib.lib()
qkv = torch.randn(B * S, 3 * d, device="cuda").to(torch.bfloat16)
out = torch.empty(B * S, d, device="cuda", dtype=torch.bfloat16)
trace = torch.zeros(H * B, 16, dtype=torch.int64, device="cuda")
for rep in range(3):
    trace.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    _lib.check(L.tamf_attn_trace(_lib.ptr(qkv), _lib.ptr(out), B, S, H, d, _lib.ptr(trace), _lib.stream_ptr()), "trace")
    e1.record()
    torch.cuda.synchronize()
print(f"B={B} S={S} d={d}: event time {e0.elapsed_time(e1) * 1e3:.1f} us for {H * B} CTAs")
t = trace.cpu().numpy()
g0 = t[:, 14].min()
names = ["start", "setup", "QK_in", "V_in", "P0_rdy", "P1_rdy", "S0_rdy", "S1_rdy", "O0_rdy", "O1_rdy", "st0", "st1", "end"]
for cta in (0, 1, 100, 147, 148, 200, 255):
    if cta >= len(t):
        continue
    r = t[cta]
    rel = " ".join(f"{n}={int(r[i] - r[0])}" for i, n in enumerate(names) if r[i])
    print(f"CTA {cta:3d} sm {int(r[13]):3d} t0 {int(r[14] - g0):6d} ns dur {int(r[15] - r[14]):6d} ns | {rel}")
dur = t[:, 15] - t[:, 14]
print("per-CTA duration ns: mean %.0f min %d max %d;  grid span %d ns" % (dur.mean(), dur.min(), dur.max(), t[:, 15].max() - g0))
starts = np.sort(t[:, 14] - g0)
print("CTA start times ns (sorted, every 16th):", starts[::16].tolist())
