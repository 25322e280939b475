"""Strided / DDIM samplers on the CUDA path (tamf_denoiser_set_sampler + the fused update epilogue) against the
reference's SpacedDiffusion.ddim_sample / p_sample outputs (tests/golden/spaced_arch_mdm.npz).  Tolerances as for the
ancestral sampler: teacher-forced step rel-L2 <= 1e-2, max-abs <= 3e-2; short free chain <= 2e-2."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from test_denoiser_gpu import ABS_TOL, REL_TOL, _dev_batch, _model

pytestmark = pytest.mark.gpu


def _setup(golden):
    from tamf_b200 import synth
    g, gs = golden("g_arch_mdm.npz"), golden("spaced_arch_mdm.npz")
    B, T = int(gs["B"]), int(gs["T"])
    m, cfg = _model("arch_mdm", g["text_feat"])
    batch = _dev_batch(synth.make_batch(B, T, nobj=3, seed=11, ragged=True))
    x = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(5)).cuda()
    return m, batch, x, gs, (B, 99, 1, T)


def test_ddim_and_strided_steps_vs_reference_golden(golden):
    import tamf_b200
    from tamf_b200 import synth
    m, batch, x, gs, shape = _setup(golden)
    B = shape[0]
    seed = int(gs["noise_seed"])
    d50 = tamf_b200.create_gaussian_diffusion(1000, "cosine", timestep_respacing="ddim50")
    mk = {"batch": batch}
    for eta in (0.0, 0.5):
        for i in (49, 20, 1, 0):
            out = d50.ddim_sample(m, x, torch.full((B,), i, dtype=torch.long, device="cuda"), clip_denoised=False,
                                  model_kwargs=mk, eta=eta, noise=synth.step_noise(seed, i, shape))
            ref = gs[f"ddim_eta{eta}_i{i}"]
            o = out["sample"].cpu().numpy()
            r, a = rel_l2(o, ref), float(np.abs(o - ref).max())
            print(f"ddim eta={eta} i={i} rel_l2={r:.3e} max_abs={a:.3e}")
            assert r <= REL_TOL and a <= ABS_TOL
            if eta == 0.0:  # the model saw the ORIGINAL timestep timestep_map[i] (respace.py:114-119)
                assert rel_l2(out["pred_xstart"].cpu().numpy(), gs[f"x0_i{i}"]) <= REL_TOL
    for i in (49, 0):  # ancestral rule on the strided process
        out = d50.p_sample(m, x, torch.full((B,), i, dtype=torch.long, device="cuda"), clip_denoised=False,
                           model_kwargs=mk, noise=synth.step_noise(seed, i, shape))
        assert rel_l2(out["sample"].cpu().numpy(), gs[f"anc_i{i}"]) <= REL_TOL
    img = x.clone()
    for i in (3, 2, 1, 0):
        img = d50.ddim_sample(m, img, torch.full((B,), i, dtype=torch.long, device="cuda"), clip_denoised=False,
                              model_kwargs=mk, eta=0.5, noise=synth.step_noise(seed, i, shape))["sample"]
    assert rel_l2(img.cpu().numpy(), gs["ddim_chain_3_0_eta0.5"]) <= 2 * REL_TOL


def test_switching_samplers_restores_the_ancestral_rule(golden):
    """ddim50 -> full 1000-step ancestral on the same model handle reproduces the ancestral golden step."""
    import tamf_b200
    from tamf_b200 import synth
    m, batch, x, gs, shape = _setup(golden)
    gp = golden("p_sample_arch_mdm.npz")
    B = shape[0]
    mk = {"batch": batch}
    d50 = tamf_b200.create_gaussian_diffusion(1000, "cosine", timestep_respacing="ddim50")
    full = tamf_b200.create_gaussian_diffusion(1000, "cosine")
    d50.ddim_sample(m, x, torch.full((B,), 49, dtype=torch.long, device="cuda"), clip_denoised=False, model_kwargs=mk)
    with pytest.raises(ValueError):  # index 500 does not exist in a 50-step process
        m.p_sample_step(x, 500, batch)
    out = full.p_sample(m, x, torch.full((B,), 500, dtype=torch.long, device="cuda"), clip_denoised=False,
                        model_kwargs=mk, noise=synth.step_noise(int(gp["noise_seed"]), 500, shape))
    assert rel_l2(out["sample"].cpu().numpy(), gp["sample_t500"]) <= REL_TOL
    # model-level forward keeps taking ORIGINAL timesteps whatever sampler is installed
    d50.ddim_sample(m, x, torch.full((B,), 49, dtype=torch.long, device="cuda"), clip_denoised=False, model_kwargs=mk)
    g = golden("g_arch_mdm.npz")
    x0 = m(x, torch.full((B,), 999, dtype=torch.long, device="cuda"), batch)
    assert rel_l2(x0.cpu().numpy(), g["x0_t999"]) <= REL_TOL


def test_ddim_loop_graph_chain():
    """ddim_sample_loop: 50 graph replays; eta = 0 is deterministic in x_T, and equals the step-by-step chain."""
    import tamf_b200
    from tamf_b200 import synth
    m, cfg = _model("arch_mdm")
    B, T = 2, 32
    batch = _dev_batch(synth.make_batch(B, T, nobj=1, seed=8))
    xT = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(2)).cuda()
    d = tamf_b200.create_gaussian_diffusion(1000, "cosine", timestep_respacing="ddim50")
    a = d.ddim_sample_loop(m, (B, 99, 1, T), noise=xT, clip_denoised=False, model_kwargs={"batch": batch}, eta=0.0)
    b = xT.clone()
    for i in range(49, -1, -1):
        b = d.ddim_sample(m, b, torch.full((B,), i, dtype=torch.long, device="cuda"), clip_denoised=False,
                          model_kwargs={"batch": batch}, eta=0.0)["sample"]
    assert torch.isfinite(a).all() and torch.equal(a, b)
    c = d.ddim_sample_loop(m, (B, 99, 1, T), noise=xT, clip_denoised=False, model_kwargs={"batch": batch}, eta=1.0,
                           seed=3)
    e = d.ddim_sample_loop(m, (B, 99, 1, T), noise=xT, clip_denoised=False, model_kwargs={"batch": batch}, eta=1.0,
                           seed=3)
    assert torch.equal(c, e) and not torch.equal(a, c)
    with pytest.raises(NotImplementedError):
        d.ddim_sample_loop(m, (B, 99, 1, T), model_kwargs={"batch": batch}, dump_steps=[1])


def test_full_ddim50_chain_vs_oracle():
    """The WHOLE 50-step DDIM chain (eta = 0, free running: every step consumes the previous CUDA result) against the
    fp32 oracle run live on the CPU with the same x_T -- end-to-end drift of the bf16 path over a complete chain."""
    import tamf_b200
    from oracle import tamf_oracle as orc
    from tamf_b200 import synth
    m, cfg = _model("arch_mdm")
    B, T = 2, 32
    batch = synth.make_batch(B, T, nobj=2, seed=8)
    xT = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(2))
    d = tamf_b200.create_gaussian_diffusion(1000, "cosine", timestep_respacing="ddim50")
    out = d.ddim_sample_loop(m, (B, 99, 1, T), noise=xT.cuda(), clip_denoised=False,
                             model_kwargs={"batch": _dev_batch(batch)}, eta=0.0).cpu()
    c1, c2, _ = d.ddim_rule(0.0)
    sd, text = synth.g_state_dict(cfg, 0), synth.text_features(batch["text"])
    x = xT.clone()
    with torch.no_grad():
        for i in range(49, -1, -1):
            ts = torch.full((B,), d.timestep_map[i], dtype=torch.long)
            x = c1[i] * orc.g_forward(sd, cfg, x, ts, batch, text) + c2[i] * x
    r = rel_l2(out.numpy(), x.numpy())
    print(f"ddim50 full chain rel_l2={r:.3e}")
    assert r <= 2 * REL_TOL


def test_ancestral_200_step_chain_vs_oracle():
    """200 free-running ancestral steps t = 199..0 with the SAME per-step noise on both sides (synth.step_noise)."""
    import tamf_b200
    from oracle import tamf_oracle as orc
    from tamf_b200 import synth
    m, cfg = _model("arch_mdm")
    B, T = 2, 24
    shape = (B, 99, 1, T)
    batch = synth.make_batch(B, T, nobj=1, seed=9)
    dbatch = _dev_batch(batch)
    x0 = torch.randn(*shape, generator=torch.Generator().manual_seed(4))
    full = tamf_b200.create_gaussian_diffusion(1000, "cosine")
    full._install(m, "ancestral")
    g = x0.cuda()
    for t in range(199, -1, -1):
        g = m.p_sample_step(g, t, dbatch, noise=synth.step_noise(31, t, shape))["sample"]
    sd, text = synth.g_state_dict(cfg, 0), synth.text_features(batch["text"])
    with torch.no_grad():
        ref = orc.p_sample_loop(sd, cfg, batch, text, shape, lambda t, s: synth.step_noise(31, t, s), x_T=x0,
                                t_start=199, t_end=0)
    r = rel_l2(g.cpu().numpy(), ref.numpy())
    print(f"ancestral 200-step chain rel_l2={r:.3e}")
    assert r <= 2 * REL_TOL
