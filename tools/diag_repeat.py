"""Debug: replay the captured one-step graph N times from the same x_t and count evaluations whose result differs from the
first one (any difference is a dependency / visibility race: the step is a pure function of its inputs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for q in (ROOT, os.path.join(ROOT, "oakink2-tamf_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, q)
import torch
import tamf_b200
from tamf_b200 import synth
from test_denoiser_gpu import _model, _dev_batch

N = int(os.environ.get("DIAG_N", "4000"))
ARCH = os.environ.get("DIAG_ARCH", "arch_mdm_l")
m, cfg = _model(ARCH)
B, T = int(os.environ.get("DIAG_B", "64")), int(os.environ.get("DIAG_T", "160"))
dbatch = _dev_batch(synth.make_batch(B, T, nobj=2, seed=4))
x = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(9)).cuda()
tamf_b200.create_gaussian_diffusion(1000, "cosine")._install(m, "ancestral")
buf = x.clone()
with m.cond_scope(dbatch, B, T, x.device) if hasattr(m, "cond_scope") else torch.no_grad():
    ref = m.p_sample_chain(buf, 519, 519, dbatch, seed=7).clone()
    bad, rows = 0, {}
    for i in range(N):
        buf.copy_(x)
        out = m.p_sample_chain(buf, 519, 519, dbatch, seed=7)
        if not torch.equal(out, ref):
            bad += 1
            d = (out - ref).abs()
            for b in sorted(set((d > 0).nonzero()[:, 0].tolist())):
                rows[b] = rows.get(b, 0) + 1
print(f"{ARCH} B={B} T={T}: evaluations {N}  differing {bad}  batch rows {rows}")
c1 = m.p_sample_chain(x.clone(), 999, 0, dbatch, seed=5).clone()
c2 = m.p_sample_chain(x.clone(), 999, 0, dbatch, seed=5).clone()
print("full 1000-step chain twice: equal", torch.equal(c1, c2), "max diff", float((c1 - c2).abs().max()))
