"""Schedule cost-constant sweep for the layer kernel (csrc/layer_chain.cuh build_layer_schedule): each combination runs
in a child process (the constants are read once per process) and times 300 graph replays of the arch_mdm_l B=64 step.
   python tools/sched_sweep.py            # parent: runs the grid, prints ms per denoiser evaluation
"""
import itertools, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oakink2-tamf_b200"))

def child():
    import torch
    import tamf_b200
    from tamf_b200 import synth
    cfg = synth.ARCH["arch_mdm_l"]
    m = tamf_b200.InterationSegmentMDM(**cfg, text_encoder=synth.text_features)
    m.load_state_dict(synth.g_state_dict(cfg, seed=0), strict=False)
    m = m.eval().cuda()
    B, T = 64, 160
    batch = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in synth.make_batch(B, T, nobj=2, seed=0).items()}
    x = torch.randn(B, 99, 1, T, device="cuda")
    tamf_b200.create_gaussian_diffusion(1000, "cosine")._install(m, "ancestral")
    m.p_sample_chain(x, 999, 900, batch, seed=1)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(2):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        m.p_sample_chain(x, 899, 600, batch, seed=1)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 300)
    print(json.dumps({"ms": best}))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    child()
    sys.exit(0)

grid = []
base = dict(TAMF_CHAIN_KB=625, TAMF_CHAIN_EPI_LN=9500, TAMF_CHAIN_LN_READY=7000, TAMF_CHAIN_EPI_GELU=5300,
            TAMF_CHAIN_SIGNAL=2500, TAMF_CHAIN_SLACK=0)
grid.append(dict(base))
for k, vals in (("TAMF_CHAIN_KB", (512, 560, 690, 760)), ("TAMF_CHAIN_SIGNAL", (0, 1200, 5000, 8000)),
                ("TAMF_CHAIN_LN_READY", (4000, 5500, 9000)), ("TAMF_CHAIN_EPI_GELU", (4300, 6000, 7000)),
                ("TAMF_CHAIN_EPI_LN", (7000, 12000)), ("TAMF_CHAIN_SLACK", (2000, 4000, 8000, 16000)),
                ("TAMF_CHAIN_EPI_BIAS", (2000, 4300))):
    for v in vals:
        g = dict(base)
        g[k] = v
        grid.append(g)
grid.append(dict(TAMF_CHAIN=0))
for g in grid:
    env = dict(os.environ, **{k: str(v) for k, v in g.items()})
    t0 = time.time()
    r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=300)
    try:
        ms = json.loads(r.stdout.strip().splitlines()[-1])["ms"]
    except Exception:
        ms = None
        print(r.stderr[-400:])
    diff = {k: v for k, v in g.items() if base.get(k) != v}
    print(f"{ms!s:>8} ms  {64 / ms / 1000 * 1000 if ms else 0:6.2f} seq/s  {diff}  ({time.time() - t0:.0f} s)", flush=True)
