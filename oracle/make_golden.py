"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the REFERENCE ITSELF (imported from
/root/reference through oracle/ref_shims.py) on seeded synthetic inputs.  Run in the authoring container:

    python -m oracle.make_golden

The fixtures travel to the GPU box (which has no /root/reference); tests compare both the oracle restatement and
the CUDA path against them.  Inputs are regenerated from the seeds recorded in each file by tamf_b200.synth, so
only reference OUTPUTS (and the CLIP-stub text features, which need the reference's CLIP code) are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oakink2-tamf_b200"))

from oracle import ref_shims  # noqa: E402
from tamf_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def golden_g(ref, arch: str, B: int, T: int, nobj: int, ragged: bool, steps, tag: str):
    cfg = synth.ARCH[arch]
    model = ref.mdm.InterationSegmentMDM(**cfg)
    model.eval()
    sd = synth.g_state_dict(cfg, seed=0)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert all(k.startswith("clip_model.") for k in missing) and not unexpected, (missing, unexpected)
    batch = synth.make_batch(B, T, nobj=nobj, seed=11, ragged=ragged)
    x = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(5))
    out = {"arch": arch, "B": B, "T": T, "nobj": nobj, "ragged": int(ragged), "batch_seed": 11, "x_seed": 5,
           "weight_seed": 0, "steps": np.array(steps)}
    with torch.no_grad():
        out["text_feat"] = model.encode_text(batch["text"]).numpy()
        for t in steps:
            ts = torch.full((B,), t, dtype=torch.long)
            out[f"x0_t{t}"] = model(x, ts, batch).numpy()
    np.savez_compressed(os.path.join(OUT, f"g_{tag}.npz"), **out)
    print("wrote g_", tag)
    return model, sd, batch, out


def golden_spaced(ref):
    """Strided / DDIM samplers of the reference (respace.py:8-111, gaussian_diffusion.py:642-690): kept timesteps,
    re-derived tables, ddim_sample / p_sample of a 50-step process at a few indices with given noise, and a free
    4-step DDIM chain.  Model = the arch_mdm reference module on the g_arch_mdm inputs."""
    cfg = synth.ARCH["arch_mdm"]
    model = ref.mdm.InterationSegmentMDM(**cfg)
    model.eval()
    model.load_state_dict(synth.g_state_dict(cfg, seed=0), strict=False)
    B, T = 3, 48
    batch = synth.make_batch(B, T, nobj=3, seed=11, ragged=True)
    x = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(5))
    gd, respace = ref.gd, ref.respace
    out = {"B": B, "T": T, "noise_seed": 77}
    for spec in ("ddim50", "ddim25", "100", "10,20,30"):
        out["steps_" + spec.replace(",", "_")] = np.array(sorted(respace.space_timesteps(1000, spec)))
    betas = gd.get_named_beta_schedule("cosine", 1000, 1.0)
    d50 = respace.SpacedDiffusion(use_timesteps=respace.space_timesteps(1000, "ddim50"), betas=betas,
                                  model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_SMALL,
                                  loss_type=gd.LossType.MSE, rescale_timesteps=False)
    out["map50"] = np.array(d50.timestep_map)
    for k in ("betas", "alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped"):
        out["d50_" + k] = getattr(d50, k)
    orig = gd.th.randn_like
    mk = {"batch": batch}
    with torch.no_grad():
        for eta in (0.0, 0.5):
            for i in (49, 20, 1, 0):
                gd.th.randn_like = lambda z, _i=i: synth.step_noise(77, _i, tuple(z.shape))
                o = d50.ddim_sample(model, x, torch.full((B,), i, dtype=torch.long), clip_denoised=False,
                                    model_kwargs=mk, eta=eta)
                out[f"ddim_eta{eta}_i{i}"] = o["sample"].numpy()
                if eta == 0.0:
                    out[f"x0_i{i}"] = o["pred_xstart"].numpy()
        for i in (49, 0):
            gd.th.randn_like = lambda z, _i=i: synth.step_noise(77, _i, tuple(z.shape))
            out[f"anc_i{i}"] = d50.p_sample(model, x, torch.full((B,), i, dtype=torch.long), clip_denoised=False,
                                            model_kwargs=mk)["sample"].numpy()
        img = x.clone()
        for i in (3, 2, 1, 0):
            gd.th.randn_like = lambda z, _i=i: synth.step_noise(77, _i, tuple(z.shape))
            img = d50.ddim_sample(model, img, torch.full((B,), i, dtype=torch.long), clip_denoised=False, model_kwargs=mk,
                                  eta=0.5)["sample"]
        out["ddim_chain_3_0_eta0.5"] = img.numpy()
    gd.th.randn_like = orig
    np.savez_compressed(os.path.join(OUT, "spaced_arch_mdm.npz"), **out)
    print("wrote spaced_arch_mdm")


def golden_g_b64(ref):
    """The BENCHMARKED configuration (BASELINE.json configs[1]: arch_mdm_l, B=64, T=160, nobj=2) through the reference
    module itself: forward at t = 999 / 0 and one p_sample with injected noise at t = 500.  Only sequences KEEP_B of the
    batch are stored (the model is batch-independent; the CUDA tile layout is not), 0.25 MB per tensor."""
    cfg = synth.ARCH["arch_mdm_l"]
    model = ref.mdm.InterationSegmentMDM(**cfg)
    model.eval()
    model.load_state_dict(synth.g_state_dict(cfg, seed=0), strict=False)
    B, T = 64, 160
    keep = np.array([0, 21, 42, 63])
    batch = synth.make_batch(B, T, nobj=2, seed=11)
    x = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(5))
    out = {"arch": "arch_mdm_l", "B": B, "T": T, "nobj": 2, "batch_seed": 11, "x_seed": 5, "weight_seed": 0,
           "keep_b": keep, "noise_seed": 77}
    diffusion = ref.diffusion_util.create_gaussian_diffusion(1000, "cosine")
    gd = ref.gd
    orig = gd.th.randn_like
    with torch.no_grad():
        for t in (999, 0):
            out[f"x0_t{t}"] = model(x, torch.full((B,), t, dtype=torch.long), batch).numpy()[keep]
        gd.th.randn_like = lambda z: synth.step_noise(77, 500, tuple(z.shape))
        o = diffusion.p_sample(model, x, torch.full((B,), 500, dtype=torch.long), clip_denoised=False,
                               model_kwargs={"batch": batch})
        out["sample_t500"] = o["sample"].numpy()[keep]
    gd.th.randn_like = orig
    np.savez_compressed(os.path.join(OUT, "g_arch_mdm_l_b64.npz"), **out)
    print("wrote g_arch_mdm_l_b64")


def golden_mano_full(ref):
    """Every MANOOutput field of the reference ManoLayer (manolayer.py:268-285) on the mano_fk.npz inputs."""
    g = np.load(os.path.join(OUT, "mano_fk.npz"))
    q, betas = torch.from_numpy(g["quat"]), torch.from_numpy(g["betas"])
    fk = {}
    for side in ("right", "left"):
        layer = ref.manolayer.ManoLayer(mano_assets_root=ref.mano_root, rot_mode="quat", side=side, center_idx=0,
                                        use_pca=False, flat_hand_mean=True)
        o = layer(pose_coeffs=q, betas=betas)
        fk[f"center_joint_{side}"] = o.center_joint.numpy()
        fk[f"transforms_abs_{side}"] = o.transforms_abs.numpy()
        fk[f"full_poses_{side}"] = o.full_poses.numpy()
    np.savez_compressed(os.path.join(OUT, "mano_fk_full.npz"), **fk)
    print("wrote mano_fk_full")


def golden_p2p(ref):
    """`point2point_signed` (model/loss/chamfer_distance.py:4-64) and `ChamferDistance.forward`
    (chamfer_distance.py:147-162) of the reference at the reference's own call shape, both directions, with normals.
    The nearest-neighbour arithmetic itself comes from the pytorch3d stub (PARITY UNPINNED, oracle/ref_shims.py);
    everything around it -- gathers, signs, norms, the 4-tuple -- is the reference's code."""
    x, xn, y, yn = synth.p2p_clouds(seed=13, T=2, nobj=2, P=8192)
    xt, xnt, yt, ynt = (torch.from_numpy(a) for a in (x, xn, y, yn))
    y2x, x2y, yidx = ref.p2p.point2point_signed(xt, yt, x_normals=xnt, y_normals=ynt)
    y2x_u, x2y_u, _ = ref.p2p.point2point_signed(xt, yt)
    cx, cy, ix, iy = ref.chd.ChamferDistance()(xt, yt)
    np.savez_compressed(os.path.join(OUT, "p2p_signed.npz"), seed=13, T=2, nobj=2, P=8192,
                        y2x_signed=y2x.numpy(), x2y_signed=x2y.numpy(), yidx_near=yidx.numpy().astype(np.int32),
                        y2x_unsigned=y2x_u.numpy(), x2y_unsigned=x2y_u.numpy(), cham_x=cx.numpy(), cham_y=cy.numpy(),
                        idx_x=ix.numpy().astype(np.int32), idx_y=iy.numpy().astype(np.int32))
    print("wrote p2p_signed")


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shims.install()
    if len(sys.argv) > 1 and sys.argv[1] == "spaced":
        return golden_spaced(ref)
    if len(sys.argv) > 1 and sys.argv[1] == "b64":
        return golden_g_b64(ref)
    if len(sys.argv) > 1 and sys.argv[1] == "mano_full":
        return golden_mano_full(ref)
    if len(sys.argv) > 1 and sys.argv[1] == "p2p":
        return golden_p2p(ref)

    # ---- G forward: arch_mdm ragged objects (pins the padded-mean quirk), 4 timesteps ----
    model, sd, batch, g = golden_g(ref, "arch_mdm", B=3, T=48, nobj=3, ragged=True, steps=[999, 500, 1, 0], tag="arch_mdm")
    # ---- G forward: arch_mdm_l at the real T ----
    golden_g(ref, "arch_mdm_l", B=2, T=160, nobj=2, ragged=False, steps=[999, 0], tag="arch_mdm_l")

    # ---- diffusion tables + p_sample chain through the reference sampler ----
    diffusion = ref.diffusion_util.create_gaussian_diffusion(1000, "cosine")
    tabs = {k: getattr(diffusion, k) for k in (
        "betas", "alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1",
        "posterior_mean_coef2", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod")}
    np.savez_compressed(os.path.join(OUT, "diffusion_tables.npz"), **tabs)
    # p_sample at a few t with given noise (reference p_sample, gaussian_diffusion.py:412-460)
    B, T = 3, 48
    x = torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(5))
    ps = {}
    gd = ref.gd
    orig = gd.th.randn_like
    for t in (999, 500, 1, 0):
        gd.th.randn_like = lambda z, _t=t: synth.step_noise(77, _t, tuple(z.shape))
        with torch.no_grad():
            o = diffusion.p_sample(model, x, torch.full((B,), t, dtype=torch.long), clip_denoised=False,
                                   model_kwargs={"batch": batch})
        ps[f"sample_t{t}"] = o["sample"].numpy()
    # free-running 4-step chain t = 3..0: the body of p_sample_loop_progressive (gaussian_diffusion.py:621-640)
    img = x.clone()
    for t in (3, 2, 1, 0):
        gd.th.randn_like = lambda z, _t=t: synth.step_noise(77, _t, tuple(z.shape))
        with torch.no_grad():
            img = diffusion.p_sample(model, img, torch.full((B,), t, dtype=torch.long), clip_denoised=False,
                                     model_kwargs={"batch": batch})["sample"]
    ps["chain_3_0"] = img.numpy()
    gd.th.randn_like = orig
    np.savez_compressed(os.path.join(OUT, "p_sample_arch_mdm.npz"), noise_seed=77, **ps)
    print("wrote p_sample")

    # ---- rotation helpers + ManoLayer FK (quat mode) on synthetic assets ----
    g0 = torch.Generator().manual_seed(3)
    N = 24
    pose_repr = torch.from_numpy(synth.random_pose_repr(np.random.default_rng(21), 1, N)[0])
    betas = 0.5 * torch.randn(N, 10, generator=g0)
    fk = {"N": N, "pose_seed": 21, "betas": betas.numpy(), "pose_repr": pose_repr.numpy()}
    R = ref.rotation.rot6d_to_rotmat(pose_repr[:, 3:].reshape(N, 16, 6))
    q = ref.rotation.rotmat_to_quat(R)
    fk["rotmat"], fk["quat"] = R.numpy(), q.numpy()
    for side in ("right", "left"):
        layer = ref.manolayer.ManoLayer(mano_assets_root=ref.mano_root, rot_mode="quat", side=side, center_idx=0,
                                        use_pca=False, flat_hand_mean=True)
        o = layer(pose_coeffs=q, betas=betas)
        fk[f"verts_{side}"], fk[f"joints_{side}"] = o.verts.numpy(), o.joints.numpy()
    np.savez_compressed(os.path.join(OUT, "mano_fk.npz"), **fk)
    print("wrote mano_fk")

    # ---- R forward (SegmentRefineModel) ----
    cfg = synth.ARCH["arch_refine"]
    rmodel = ref.refine.SegmentRefineModel(ref.mano_root, **cfg, use_pc=True)
    rmodel.eval()
    rsd = synth.r_state_dict(cfg, 0)
    missing, unexpected = rmodel.load_state_dict(rsd, strict=False)
    assert all("mano_layer" in k for k in missing) and not unexpected, (missing, unexpected)
    B, T, P = 2, 16, 256
    rb = synth.make_batch(B, T, nobj=2, seed=2, ragged=True, npoints=P, with_pointcloud=True)
    with torch.no_grad():
        ro = rmodel(rb)
    keep = ["refine_pose_repr", "sample_hand_verts", "sample_hand_joints", "sample_h2o_dist", "refine_h2o_dist",
            "target_h2o_dist", "refine_hand_verts", "target_hand_joints"]
    np.savez_compressed(os.path.join(OUT, "r_arch_refine.npz"), B=B, T=T, P=P, batch_seed=2, weight_seed=0,
                        **{k: ro[k].numpy() for k in keep})
    print("wrote r_arch_refine")
    golden_spaced(ref)
    golden_mano_full(ref)
    golden_g_b64(ref)
    golden_p2p(ref)


if __name__ == "__main__":
    main()
