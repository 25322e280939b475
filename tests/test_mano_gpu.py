"""ManoLayer FK parity: golden vectors produced by the reference's own manotorch code + live oracle, 1e-5 m."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-5  # metres (north star)


@pytest.mark.parametrize("side", ["right", "left"])
def test_fk_quat_vs_reference_golden(golden, side):
    import tamf_b200
    from tamf_b200 import synth
    g = golden("mano_fk.npz")
    layer = tamf_b200.ManoLayer(rot_mode="quat", side=side, center_idx=0, use_pca=False, flat_hand_mean=True,
                                assets=synth.mano_assets(side))
    out = layer(pose_coeffs=torch.from_numpy(g["quat"]).cuda(), betas=torch.from_numpy(g["betas"]).cuda())
    assert out.verts.shape == (g["quat"].shape[0], 778, 3) and out.joints.shape == (g["quat"].shape[0], 21, 3)
    assert np.abs(out.verts.cpu().numpy() - g[f"verts_{side}"]).max() < TOL
    assert np.abs(out.joints.cpu().numpy() - g[f"joints_{side}"]).max() < TOL
    assert float(out.joints[:, 0].abs().max()) == 0.0  # root-centred exactly (center_idx = 0)


@pytest.mark.parametrize("side", ["right", "left"])
def test_mano_output_fields_vs_reference_golden(golden, side):
    """Every MANOOutput field against the reference ManoLayer (manolayer.py:242-285): `center_joint` is the root joint
    BEFORE centring (verts + center_joint gives the uncentred mesh), `transforms_abs` the centre-shifted global joint
    transforms, `full_poses` the axis-angle form of the input quaternions."""
    import tamf_b200
    from tamf_b200 import synth
    g, gf = golden("mano_fk.npz"), golden("mano_fk_full.npz")
    layer = tamf_b200.ManoLayer(rot_mode="quat", side=side, center_idx=0, use_pca=False, flat_hand_mean=True,
                                assets=synth.mano_assets(side))
    betas = torch.from_numpy(g["betas"]).cuda()
    out = layer(pose_coeffs=torch.from_numpy(g["quat"]).cuda(), betas=betas)
    N = g["quat"].shape[0]
    assert out.center_joint.shape == (N, 1, 3) and out.transforms_abs.shape == (N, 16, 4, 4)
    assert out.full_poses.shape == (N, 48) and out.center_idx == 0 and out.betas is betas
    assert np.abs(out.center_joint.cpu().numpy() - gf[f"center_joint_{side}"]).max() < TOL
    assert float(out.center_joint.abs().max()) > 1e-3  # not the zeros of round 1
    assert np.abs(out.transforms_abs.cpu().numpy() - gf[f"transforms_abs_{side}"]).max() < TOL
    assert np.abs(out.full_poses.cpu().numpy() - gf[f"full_poses_{side}"]).max() < 1e-5
    # joints 0..15 of the kinematic chain are the translations of transforms_abs (reordered, manolayer.py:240)
    order = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]
    chain = {slot: k for slot, k in enumerate(order) if k < 16}
    for slot, k in chain.items():
        assert torch.equal(out.joints[:, slot], out.transforms_abs[:, k, :3, 3])


@pytest.mark.parametrize("side", ["right", "left"])
@pytest.mark.parametrize("N", [1, 7, 8, 9, 160, 1000])
def test_fk_pose_repr_vs_oracle(side, N):
    """pose_repr front end (rot6d -> rotmat -> quat -> FK -> + tsl) incl. ragged tail tiles (N % 8 != 0)."""
    import tamf_b200
    from oracle import tamf_oracle as orc
    from tamf_b200 import synth
    A = synth.mano_assets(side)
    layer = tamf_b200.ManoLayer(side=side, assets=A)
    rng = np.random.default_rng(N)
    pose = torch.from_numpy(synth.random_pose_repr(rng, 1, N)[0])
    betas = torch.from_numpy((0.5 * rng.standard_normal((N, 10))).astype(np.float32))
    v, j = layer.forward_pose_repr(pose.cuda(), betas.cuda())
    rv, rj = orc.mano_fk_pose_repr(A, pose, betas, side)
    assert np.abs(v.cpu().numpy() - rv.numpy()).max() < TOL
    assert np.abs(j.cpu().numpy() - rj.numpy()).max() < TOL


def test_fk_full_size_vs_oracle():
    """BASELINE size: all 64 x 160 = 10 240 frames of a batch, both hand sides, against the live oracle, 1e-5 m."""
    import tamf_b200
    from oracle import tamf_oracle as orc
    from tamf_b200 import synth
    N = 64 * 160
    rng = np.random.default_rng(77)
    pose = torch.from_numpy(synth.random_pose_repr(rng, 1, N)[0])
    betas = torch.from_numpy((0.5 * rng.standard_normal((N, 10))).astype(np.float32))
    for side in ("right", "left"):
        A = synth.mano_assets(side)
        layer = tamf_b200.ManoLayer(side=side, assets=A)
        v, j = layer.forward_pose_repr(pose.cuda(), betas.cuda())
        rv, rj = orc.mano_fk_pose_repr(A, pose, betas, side)
        ev, ej = float((v.cpu() - rv).abs().max()), float((j.cpu() - rj).abs().max())
        print(f"FK {side} N={N}: max |dv| {ev:.2e} m, max |dj| {ej:.2e} m")
        assert ev < TOL and ej < TOL


def test_fk_full_size_properties():
    """BASELINE size (B=64 x T=160 frames): translation equivariance and frame independence."""
    import tamf_b200
    from tamf_b200 import synth
    layer = tamf_b200.ManoLayer(side="right", assets=synth.mano_assets("right"))
    N = 64 * 160
    rng = np.random.default_rng(1)
    pose = torch.from_numpy(synth.random_pose_repr(rng, 1, N)[0]).cuda()
    betas = torch.from_numpy((0.5 * rng.standard_normal((N, 10))).astype(np.float32)).cuda()
    v, j = layer.forward_pose_repr(pose, betas)
    shift = torch.tensor([0.25, -0.5, 1.0], device="cuda")
    pose2 = pose.clone()
    pose2[:, :3] += shift
    v2, j2 = layer.forward_pose_repr(pose2, betas)
    assert float((v2 - v - shift).abs().max()) < 2e-6
    perm = torch.randperm(N, device="cuda")
    v3, _ = layer.forward_pose_repr(pose[perm], betas[perm])
    assert torch.equal(v3, v[perm])


def test_fk_errors():
    import tamf_b200
    from tamf_b200 import synth
    layer = tamf_b200.ManoLayer(side="right", assets=synth.mano_assets("right"))
    with pytest.raises(ValueError):
        layer(pose_coeffs=torch.zeros(4, 15, 4, device="cuda"), betas=torch.zeros(4, 10, device="cuda"))
    with pytest.raises(NotImplementedError):
        tamf_b200.ManoLayer(rot_mode="axisang", assets=synth.mano_assets("right"))
