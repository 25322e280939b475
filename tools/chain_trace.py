"""Debug aid: per-CTA timeline (clock64) and CUDA-event time of ONE chain-kernel launch (csrc/gemm_chain.cuh) at the
benchmarked shapes.  Run under gpurun:   python tools/chain_trace.py [which] [M] [d] [ff]
   which 0: out_proj + LN1 -> linear1 + GELU (kernel A), 1: linear2 + LN2 -> next in_proj (kernel B), 2: linear2 + LN2 only"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oakink2-tamf_b200"))
import torch
from tamf_b200 import _lib

which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
M = int(sys.argv[2]) if len(sys.argv) > 2 else 10560
d = int(sys.argv[3]) if len(sys.argv) > 3 else 512
ff = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
K1 = d if which == 0 else ff
N2 = {0: ff, 1: 3 * d, 2: 0}[which]
L = _lib.lib()
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
a1 = rn(M, K1).to(torch.bfloat16)
w1 = (rn(d, K1) / K1 ** 0.5).to(torch.bfloat16)
lnp = torch.cat([0.1 * rn(d), 1 + 0.1 * rn(d), 0.1 * rn(d)]).contiguous()
x = rn(M, d)
w2 = (rn(max(N2, 256), d) / d ** 0.5).to(torch.bfloat16)
b2 = 0.1 * rn(max(N2, 256))
c2 = torch.empty(M, max(N2, 256), device="cuda", dtype=torch.bfloat16)
nb = L.tamf_chain_aux_bytes(M, d, max(K1, N2, 3 * d))
aux = torch.zeros(nb, dtype=torch.uint8, device="cuda")
trace = torch.zeros(148, 64, dtype=torch.int64, device="cuda")
times = []
for rep in range(5):
    xh = x.to(torch.bfloat16)
    xl = (x - xh.float()).to(torch.bfloat16)
    trace.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    _lib.check(L.tamf_chain_run(which, _lib.ptr(a1), _lib.ptr(w1), _lib.ptr(lnp), _lib.ptr(xh), _lib.ptr(xl), _lib.ptr(w2),
                                _lib.ptr(b2), _lib.ptr(c2), M, d, K1, N2, _lib.ptr(aux), nb, _lib.ptr(trace),
                                _lib.stream_ptr()), "chain_run")
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) * 1e3)
fl = 2 * M * d * K1 + 2 * M * N2 * d
print(f"which={which} M={M} d={d} K1={K1} N2={N2}: event time (incl. schedule upload) {min(times):.1f} us, {fl/1e9:.1f} GFLOP")
t = trace.cpu().numpy()
ends = [int(r[3] - r[0]) for r in t if r[0]]
print(f"CTAs {len(ends)}  kernel span per CTA (cycles): min {min(ends)} max {max(ends)}")
for cta in (0, 1, 2, 18, 19, 20, 21, 72, 73, 146, 147):
    r = t[cta]
    if r[0] == 0:
        continue
    z = r[0]
    f = lambda v: "   -  " if v == 0 else f"{(v - z):6d}"
    print(f"CTA {cta}: setup {f(r[1])} pdl {f(r[2])} end {f(r[3])}  first-LN stats done {f(r[58])}")
    for it in range(9):
        if r[4 + 6 * it] == 0 and r[8 + 6 * it] == 0:
            break
        print(f"   unit {it}: prod [{f(r[4+6*it])},{f(r[5+6*it])}]  mma [{f(r[6+6*it])},{f(r[7+6*it])}]  epi [{f(r[8+6*it])},{f(r[9+6*it])}]")
