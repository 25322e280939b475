"""Probe: does running the batch as K independent sub-batch chains on K streams fill the SMs the LayerNorm GEMMs
(42 CTA pairs of 74) and the kernel ramps/tails leave idle?  Usage: python tools/dual_stream_probe.py [chain_steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oakink2-tamf_b200")]
import torch  # noqa: E402

import tamf_b200  # noqa: E402
from tamf_b200 import synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda:0")
cfg = synth.ARCH["arch_mdm_l"]
sd = synth.g_state_dict(cfg, seed=0)
T = 160


def make(B, seed):
    m = tamf_b200.InterationSegmentMDM(**cfg, text_encoder=synth.text_features)
    m.load_state_dict(sd, strict=False)
    m = m.eval().to(dev)
    b = synth.make_batch(B, T, nobj=2, seed=seed)
    b = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in b.items()}
    x = torch.randn(B, 99, 1, T, device=dev)
    return m, b, x


def run(parts, label):
    streams = [torch.cuda.Stream(dev) for _ in parts]
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        for s in streams:
            s.wait_stream(torch.cuda.current_stream())
        # interleave chunks so neither stream's host-side queue starves
        chunk = 25
        for c0 in range(0, steps, chunk):
            for (m, b, x), s in zip(parts, streams):
                with torch.cuda.stream(s):
                    m.p_sample_chain(x, 999 - c0, 999 - min(steps, c0 + chunk) + 1, b, seed=5)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    nseq = sum(p[2].shape[0] for p in parts)
    print(f"{label}: {ms / steps:.4f} ms per denoiser evaluation of {nseq} sequences -> "
          f"{nseq / (ms / steps * 1e-3 * 1000):.2f} seq/s for a 1000-step chain", flush=True)


run([make(64, 1)], "1 x B=64")
run([make(32, 1)], "1 x B=32")
run([make(32, 1), make(32, 2)], "2 x B=32 on 2 streams")
run([make(16, i) for i in range(4)], "4 x B=16 on 4 streams")
run([make(64, 1), make(64, 2)], "2 x B=64 on 2 streams")
