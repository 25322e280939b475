"""Debug aid: per-CTA timeline (clock64) of one hot-path GEMM launch.  Run under gpurun:
   python tools/gemm_trace.py [which]      which: 0 in_proj, 1 linear1+GELU, 2 LN GEMM (linear2 shape), 3 LN GEMM (out_proj shape),
   4 embed-b (token epilogue), 5 embed-a (SiLU)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oakink2-tamf_b200"))
import torch
from tamf_b200 import _lib

which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
M = 10240 if which in (4, 5) else 10560
N, K = {0: (1536, 512), 1: (2048, 512), 2: (512, 2048), 3: (512, 512), 4: (512, 512), 5: (512, 128), 10: (1536, 512)}[which]
w_id = 2 if which == 3 else which  # 4: embed-b (token epilogue), 5: embed-a (SiLU)
L = _lib.lib()
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
out = torch.empty(10560, N, device="cuda", dtype=torch.bfloat16)
X = torch.randn(10560, N, device="cuda")
trace = torch.zeros(148, 64, dtype=torch.int64, device="cuda")
for rep in range(3):
    trace.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    _lib.check(L.tamf_gemm_trace(w_id, _lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), _lib.ptr(X), M, N, K,
                                 _lib.ptr(trace), _lib.stream_ptr()), "trace")
    e1.record()
    torch.cuda.synchronize()
print(f"which={which} M={M} N={N} K={K}  event time {e0.elapsed_time(e1)*1e3:.1f} us")
t = trace.cpu().numpy()
for cta in (0, 1, 73, 147):
    r = t[cta]
    if r[0] == 0:
        continue
    z = r[0]
    f = lambda v: "   -  " if v == 0 else f"{(v - z):6d}"
    print(f"CTA {cta}: setup_done {f(r[1])} pdl_wait_done {f(r[2])} end {f(r[3])}  ln: pass1_done {f(r[4])} stats_bar {f(r[5])} pass2_done {f(r[6])}")
    for it in range(8):
        if r[8 + 2 * it] == 0:
            break
        print(f"   tile {it}: prod [{f(r[8+2*it])},{f(r[9+2*it])}]  mma [{f(r[24+2*it])},{f(r[25+2*it])}]  epi [{f(r[40+2*it])},{f(r[41+2*it])}]")
    print("   mma tile1 full-acquired per k-block:", [f(r[k]) for k in range(56, 64)])
