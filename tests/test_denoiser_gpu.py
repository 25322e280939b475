"""MF-MDM G denoiser + DDPM posterior parity.  Tolerances (north star / SURVEY.md 8d): per-step x0 and x_{t-1}
under teacher forcing, bf16 tensor-core path vs the fp32 reference: rel-L2 <= 1e-2, max-abs <= 3e-2."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
REL_TOL, ABS_TOL = 1e-2, 3e-2


def _model(arch, text_feat=None):
    import tamf_b200
    from tamf_b200 import synth
    cfg = synth.ARCH[arch]
    enc = (lambda texts: torch.from_numpy(text_feat)) if text_feat is not None else synth.text_features
    m = tamf_b200.InterationSegmentMDM(**cfg, text_encoder=enc)
    missing, unexpected = m.load_state_dict(synth.g_state_dict(cfg, seed=0), strict=False)
    assert not missing and not unexpected
    m.eval()
    return m.to("cuda"), cfg


def _dev_batch(batch):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


@pytest.mark.parametrize("tag", ["arch_mdm", "arch_mdm_l"])
def test_forward_vs_reference_golden(golden, tag):
    """x0 = model(x_t, t, batch) against outputs of the reference's own InterationSegmentMDM (tests/golden)."""
    from tamf_b200 import synth
    g = golden(f"g_{tag}.npz")
    B, T = int(g["B"]), int(g["T"])
    m, cfg = _model(str(g["arch"]), g["text_feat"])
    batch = _dev_batch(synth.make_batch(B, T, nobj=int(g["nobj"]), seed=int(g["batch_seed"]), ragged=bool(g["ragged"])))
    x = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(int(g["x_seed"]))).cuda()
    for t in g["steps"]:
        out = m(x, torch.full((B,), int(t), dtype=torch.long, device="cuda"), batch)
        ref = g[f"x0_t{int(t)}"]
        o = out.cpu().numpy()
        assert o.shape == ref.shape
        r, a = rel_l2(o, ref), float(np.abs(o - ref).max())
        print(f"{tag} t={int(t)} rel_l2={r:.3e} max_abs={a:.3e}")
        assert r <= REL_TOL and a <= ABS_TOL


def test_forward_per_row_timesteps_vs_oracle():
    """forward() accepts a different t per row (training-style call), checked against the live oracle."""
    from oracle import tamf_oracle as orc
    from tamf_b200 import synth
    m, cfg = _model("arch_mdm")
    B, T = 4, 40
    batch = synth.make_batch(B, T, nobj=2, seed=3)
    x = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(1))
    ts = torch.tensor([999, 0, 17, 500])
    ref = orc.g_forward(synth.g_state_dict(cfg, 0), cfg, x, ts, batch, synth.text_features(batch["text"]))
    out = m(x.cuda(), ts.cuda(), _dev_batch(batch)).cpu()
    assert rel_l2(out.numpy(), ref.numpy()) <= REL_TOL


def test_p_sample_vs_reference_golden(golden):
    """One ancestral step with the reference's own noise (teacher forcing) at t in {999,500,1,0} and the
    free-running 4-step chain t=3..0, against GaussianDiffusion.p_sample outputs of the reference."""
    import tamf_b200
    from tamf_b200 import synth
    g, gp = golden("g_arch_mdm.npz"), golden("p_sample_arch_mdm.npz")
    B, T = int(g["B"]), int(g["T"])
    m, cfg = _model("arch_mdm", g["text_feat"])
    batch = _dev_batch(synth.make_batch(B, T, nobj=int(g["nobj"]), seed=int(g["batch_seed"]), ragged=True))
    x = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(5)).cuda()
    diffusion = tamf_b200.create_gaussian_diffusion(1000, "cosine")
    seed = int(gp["noise_seed"])
    for t in (999, 500, 1, 0):
        out = diffusion.p_sample(m, x, torch.full((B,), t, dtype=torch.long, device="cuda"), clip_denoised=False,
                                 model_kwargs={"batch": batch}, noise=synth.step_noise(seed, t, (B, 99, 1, T)))
        ref = gp[f"sample_t{t}"]
        o = out["sample"].cpu().numpy()
        r, a = rel_l2(o, ref), float(np.abs(o - ref).max())
        print(f"p_sample t={t} rel_l2={r:.3e} max_abs={a:.3e}")
        assert r <= REL_TOL and a <= ABS_TOL
    img = x.clone()
    for t in (3, 2, 1, 0):
        img = m.p_sample_step(img, t, batch, noise=synth.step_noise(seed, t, (B, 99, 1, T)))["sample"]
    assert rel_l2(img.cpu().numpy(), gp["chain_3_0"]) <= 2 * REL_TOL


def test_chain_graph_matches_stepwise():
    """CUDA-graph chain with in-kernel Philox == the same steps issued one by one with the same (seed, t)."""
    from tamf_b200 import synth
    m, cfg = _model("arch_mdm")
    B, T = 2, 32
    batch = _dev_batch(synth.make_batch(B, T, nobj=1, seed=8))
    x = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(2)).cuda()
    a = m.p_sample_chain(x.clone(), 9, 0, batch, seed=1234)
    b = x.clone()
    for t in range(9, -1, -1):
        b = m.p_sample_step(b, t, batch, noise=None, seed=1234)["sample"]
    assert torch.equal(a, b)
    c = m.p_sample_chain(x.clone(), 9, 0, batch, seed=1234)
    assert torch.equal(a, c)  # deterministic replay


def test_full_1000_step_chain_two_hand_sequence():
    """BASELINE.json configs[0] as a parity test: arch_mdm, ONE two-hand sequence (rows rh + lh sharing text and
    objects), T=160, nobj=2, the full 1000-step ancestral chain, free running on both sides with the same per-step
    noise: CUDA bf16 path vs the fp32 oracle on the host cores (about 10 s).  tools/parity_full_chain.py prints the
    drift along the chain (profiles/r01_parity_full_chain.json: 3.7e-3 at t = 0)."""
    import tamf_b200
    from oracle import tamf_oracle as orc
    from tamf_b200 import synth
    m, cfg = _model("arch_mdm")
    B, T = 2, 160
    batch = synth.make_batch(B, T, nobj=2, seed=0)
    for k in ("text", "obj_list"):
        if k in batch:
            batch[k] = [batch[k][0]] * B
    for k in ("shape", "obj_traj", "obj_embedding"):
        batch[k] = batch[k][:1].repeat(B, *([1] * (batch[k].ndim - 1)))
    batch["hand_side"] = ["rh", "lh"]
    shape = (B, 99, 1, T)
    sd, text, tab = synth.g_state_dict(cfg, 0), synth.text_features(batch["text"]), orc.diffusion_tables(1000)
    tamf_b200.create_gaussian_diffusion(1000, "cosine")._install(m, "ancestral")
    dbatch = _dev_batch(batch)
    xT = synth.step_noise(123, 1000, shape)
    g, r = xT.cuda(), xT.clone()
    with torch.no_grad():
        for t in range(999, -1, -1):
            n = synth.step_noise(123, t, shape)
            g = m.p_sample_step(g, t, dbatch, noise=n)["sample"]
            r = orc.p_sample_update(tab, r, orc.g_forward(sd, cfg, r, torch.full((B,), t, dtype=torch.long), batch, text),
                                    t, n)
    rl = rel_l2(g.cpu().numpy(), r.numpy())
    print(f"full 1000-step chain rel_l2={rl:.3e}")
    assert torch.isfinite(g).all() and rl <= 2 * REL_TOL
    assert not torch.equal(g[0], g[1])  # rh / lh rows differ only through the hand-side token


def test_philox_normal_statistics():
    from tamf_b200 import _lib
    n = 1 << 22
    out = torch.empty(n, device="cuda")
    _lib.check(_lib.lib().tamf_philox_normal(_lib.ptr(out), n, 42, 7, _lib.stream_ptr()), "philox")
    assert abs(out.mean().item()) < 3e-3 and abs(out.std().item() - 1.0) < 3e-3
    assert abs((out ** 4).mean().item() - 3.0) < 0.05
    out2 = torch.empty(n, device="cuda")
    _lib.check(_lib.lib().tamf_philox_normal(_lib.ptr(out2), n, 42, 8, _lib.stream_ptr()), "philox")
    assert abs(torch.corrcoef(torch.stack([out, out2]))[0, 1].item()) < 3e-3


def test_b64_forward_and_p_sample_vs_reference_golden(golden):
    """The BENCHMARKED configuration (BASELINE.json configs[1]: arch_mdm_l, B=64, T=160, nobj=2; M = 10 560 token rows
    = 42 pair tiles, several tiles per persistent CTA pair) against outputs of the reference's own module
    (tests/golden/g_arch_mdm_l_b64.npz, sequences keep_b of the batch): forward at t = 999 / 0, p_sample at t = 500."""
    import tamf_b200
    from tamf_b200 import synth
    g = golden("g_arch_mdm_l_b64.npz")
    B, T, keep = int(g["B"]), int(g["T"]), g["keep_b"]
    m, cfg = _model(str(g["arch"]))
    batch = _dev_batch(synth.make_batch(B, T, nobj=int(g["nobj"]), seed=int(g["batch_seed"])))
    x = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(int(g["x_seed"]))).cuda()
    for t in (999, 0):
        o = m(x, torch.full((B,), t, dtype=torch.long, device="cuda"), batch).cpu().numpy()[keep]
        ref = g[f"x0_t{t}"]
        r, a = rel_l2(o, ref), float(np.abs(o - ref).max())
        print(f"B=64 forward t={t} rel_l2={r:.3e} max_abs={a:.3e}")
        assert r <= REL_TOL and a <= ABS_TOL
    diffusion = tamf_b200.create_gaussian_diffusion(1000, "cosine")
    out = diffusion.p_sample(m, x, torch.full((B,), 500, dtype=torch.long, device="cuda"), clip_denoised=False,
                             model_kwargs={"batch": batch},
                             noise=synth.step_noise(int(g["noise_seed"]), 500, (B, 99, 1, T)))
    o, ref = out["sample"].cpu().numpy()[keep], g["sample_t500"]
    r, a = rel_l2(o, ref), float(np.abs(o - ref).max())
    print(f"B=64 p_sample t=500 rel_l2={r:.3e} max_abs={a:.3e}")
    assert r <= REL_TOL and a <= ABS_TOL


def test_b64_forward_and_chain_vs_live_oracle():
    """Same configuration against the LIVE oracle over the whole batch: (a) one forward, all 64 sequences; (b) a 20-step
    CUDA-graph chain (in-kernel Philox) == the same 20 steps issued one by one, bit for bit; (c) the oracle run free over
    those 20 steps with the noise the kernels drew (recovered from x_{t-1} = c1 x0 + c2 x_t + sigma eps) lands on the
    graph chain's result: rel-L2 <= 1e-2, max-abs <= 3e-2."""
    import tamf_b200
    from oracle import tamf_oracle as orc
    from tamf_b200 import synth
    m, cfg = _model("arch_mdm_l")
    B, T = 64, 160
    batch = synth.make_batch(B, T, nobj=2, seed=4)
    dbatch = _dev_batch(batch)
    sd, text, tab = synth.g_state_dict(cfg, 0), synth.text_features(batch["text"]), orc.diffusion_tables(1000)
    x = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(9))
    ts = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = orc.g_forward(sd, cfg, x, ts, batch, text)
    out = m(x.cuda(), ts.cuda(), dbatch).cpu()
    r, a = rel_l2(out.numpy(), ref.numpy()), float((out - ref).abs().max())
    print(f"B=64 forward vs oracle rel_l2={r:.3e} max_abs={a:.3e}")
    assert r <= REL_TOL and a <= ABS_TOL
    # (a') one p_sample with injected noise against the oracle's update of the oracle's own x0 (gaussian_diffusion.py:412-460)
    tamf_b200.create_gaussian_diffusion(1000, "cosine")._install(m, "ancestral")
    tq = 377
    eps_in = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        x0_ref = orc.g_forward(sd, cfg, x, torch.full((B,), tq, dtype=torch.long), batch, text)
        want = orc.p_sample_update(tab, x, x0_ref, tq, eps_in)
    got = m.p_sample_step(x.cuda().clone(), tq, dbatch, noise=eps_in.cuda())
    r, a = rel_l2(got["sample"].cpu().numpy(), want.numpy()), float((got["sample"].cpu() - want).abs().max())
    print(f"B=64 p_sample (injected noise) vs oracle rel_l2={r:.3e} max_abs={a:.3e}")
    assert r <= REL_TOL and a <= ABS_TOL
    assert rel_l2(got["pred_xstart"].cpu().numpy(), x0_ref.numpy()) <= REL_TOL
    # (b) graph chain == stepwise
    t0, t1, seed = 519, 500, 2024
    chain = m.p_sample_chain(x.cuda().clone(), t0, t1, dbatch, seed=seed)
    cur, eps = x.cuda().clone(), {}
    for t in range(t0, t1 - 1, -1):
        o = m.p_sample_step(cur, t, dbatch, noise=None, seed=seed)
        c1, c2 = float(np.float32(tab["posterior_mean_coef1"][t])), float(np.float32(tab["posterior_mean_coef2"][t]))
        sg = float(np.exp(np.float32(0.5) * np.float32(tab["posterior_log_variance_clipped"][t])))
        eps[t] = ((o["sample"].double() - c1 * o["pred_xstart"].double() - c2 * cur.double()) / sg).float().cpu()
        cur = o["sample"]
    assert torch.equal(chain, cur)
    e = torch.stack(list(eps.values()))
    assert abs(float(e.mean())) < 5e-3 and abs(float(e.std()) - 1.0) < 5e-3  # the recovered noise is standard normal
    # (c) oracle, free running on the same noise
    r_x = x.clone()
    with torch.no_grad():
        for t in range(t0, t1 - 1, -1):
            x0 = orc.g_forward(sd, cfg, r_x, torch.full((B,), t, dtype=torch.long), batch, text)
            r_x = orc.p_sample_update(tab, r_x, x0, t, eps[t])
    r, a = rel_l2(chain.cpu().numpy(), r_x.numpy()), float((chain.cpu() - r_x).abs().max())
    print(f"B=64 20-step graph chain vs oracle rel_l2={r:.3e} max_abs={a:.3e}")
    assert r <= REL_TOL and a <= ABS_TOL


def test_step_graph_replays_are_bit_identical():
    """The captured step is a pure function of (x_t, t, seed).  Every unit of the layer kernel and every attention CTA
    starts on per-row-tile / per-sequence counters instead of kernel boundaries (layer_chain.cuh), so a missing release or
    proxy fence shows up as an occasional evaluation that read stale rows: with relaxed counter updates 15 of 4000 replays
    at this shape differed in one sequence (~1e-4).  1500 replays, all equal to the first, bit for bit."""
    import tamf_b200
    from tamf_b200 import synth
    m, cfg = _model("arch_mdm_l")
    B, T = 64, 160
    dbatch = _dev_batch(synth.make_batch(B, T, nobj=2, seed=4))
    x = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(9)).cuda()
    tamf_b200.create_gaussian_diffusion(1000, "cosine")._install(m, "ancestral")
    buf = x.clone()
    with m.cond_scope(dbatch, B, T, x.device):
        ref = m.p_sample_chain(buf, 519, 519, dbatch, seed=7).clone()
        differing = 0
        for _ in range(1500):
            buf.copy_(x)
            differing += int(not torch.equal(m.p_sample_chain(buf, 519, 519, dbatch, seed=7), ref))
    assert differing == 0


def test_full_size_forward_properties():
    """BASELINE size (arch_mdm_l, B=64, T=160): batch-row independence (each chain depends only on its own row) and
    agreement of row 0 with a B=1 evaluation -- size-independent properties, no oracle needed."""
    from tamf_b200 import synth
    m, cfg = _model("arch_mdm_l")
    B, T = 64, 160
    batch = synth.make_batch(B, T, nobj=2, seed=0)
    x = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(0)).cuda()
    ts = torch.full((B,), 321, dtype=torch.long, device="cuda")
    full = m(x, ts, _dev_batch(batch))
    assert torch.isfinite(full).all()
    sub = {k: (v[:3] if isinstance(v, (torch.Tensor, list)) else v) for k, v in batch.items()}
    part = m(x[:3], ts[:3], _dev_batch(sub))
    assert rel_l2(part.cpu().numpy(), full[:3].cpu().numpy()) < 1e-5


def test_errors():
    from tamf_b200 import synth
    m, cfg = _model("arch_mdm")
    batch = _dev_batch(synth.make_batch(2, 16, nobj=1, seed=0))
    x = torch.zeros(2, 99, 1, 16, device="cuda")
    bad = dict(batch)
    bad["hand_side"] = ["rh", "xx"]
    with pytest.raises(ValueError, match="unexpected hand_side"):
        m(x, torch.zeros(2, dtype=torch.long, device="cuda"), bad)
    with pytest.raises(ValueError):
        m(torch.zeros(2, 98, 1, 16, device="cuda"), torch.zeros(2, dtype=torch.long, device="cuda"), batch)


@pytest.mark.parametrize("B,T,nobj", [(1, 171, 1), (3, 1, 2), (2, 7, 8), (5, 123, 3), (3, 64, 2), (5, 96, 1), (7, 32, 2)])
def test_forward_edge_shapes_vs_oracle(B, T, nobj):
    """Edge shapes against the live fp32 oracle: the longest sequence the path accepts (T = 171: 176 tokens, the full
    attention tile), a single frame, the maximum object count (8, ragged + zero padded), a frame count that is not a
    multiple of any tile size, and frame counts that are multiples of 32 (TMA-store form of the token epilogue, with
    row tiles that end inside and past the batch)."""
    from oracle import tamf_oracle as orc
    from tamf_b200 import synth
    m, cfg = _model("arch_mdm")
    batch = synth.make_batch(B, T, nobj=nobj, seed=B * 100 + T, ragged=nobj > 1)
    x = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(T))
    ts = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(B))
    with torch.no_grad():
        ref = orc.g_forward(synth.g_state_dict(cfg, 0), cfg, x, ts, batch, synth.text_features(batch["text"]))
    out = m(x.cuda(), ts.cuda(), _dev_batch(batch)).cpu()
    r, a = rel_l2(out.numpy(), ref.numpy()), float((out - ref).abs().max())
    print(f"B={B} T={T} nobj={nobj} rel_l2={r:.3e} max_abs={a:.3e}")
    assert r <= REL_TOL and a <= ABS_TOL


def test_sequence_too_long_is_rejected():
    from tamf_b200 import synth
    m, cfg = _model("arch_mdm")
    batch = _dev_batch(synth.make_batch(1, 172, nobj=1, seed=0))
    with pytest.raises(ValueError, match="176"):
        m(torch.zeros(1, 99, 1, 172, device="cuda"), torch.zeros(1, dtype=torch.long, device="cuda"), batch)
