"""MF-MDM R (SegmentRefineModel) parity: golden outputs of the reference's own module + live oracle."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _model():
    import tamf_b200
    from tamf_b200 import synth
    cfg = synth.ARCH["arch_refine"]
    m = tamf_b200.SegmentRefineModel("unused", **cfg, use_pc=True,
                                     mano_assets={"right": synth.mano_assets("right"), "left": synth.mano_assets("left")})
    missing, unexpected = m.load_state_dict(synth.r_state_dict(cfg, 0), strict=False)
    assert not unexpected and all("mano_layer" in k for k in missing), (missing, unexpected)
    return m.eval().to("cuda"), cfg


def _dev(batch):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


def test_r_forward_vs_reference_golden(golden):
    from tamf_b200 import synth
    g = golden("r_arch_refine.npz")
    B, T, P = int(g["B"]), int(g["T"]), int(g["P"])
    m, cfg = _model()
    batch = synth.make_batch(B, T, nobj=2, seed=int(g["batch_seed"]), ragged=True, npoints=P, with_pointcloud=True)
    out = m(_dev(batch))
    assert set(out.keys()) == {
        "refine_pose_repr", "refine_hand_verts", "refine_hand_joints", "refine_hand_normals", "refine_h2o_dist",
        "target_hand_verts", "target_hand_joints", "target_hand_normals", "target_h2o_dist", "sample_hand_verts",
        "sample_hand_joints", "sample_hand_normals", "sample_h2o_dist"}
    for k in ("sample_hand_verts", "sample_hand_joints", "target_hand_joints"):
        assert np.abs(out[k].cpu().numpy() - g[k]).max() < 1e-5, k  # metres (north star)
    for k in ("sample_h2o_dist", "target_h2o_dist"):
        assert np.abs(out[k].cpu().numpy() - g[k]).max() < 2e-5, k
    # bf16 transformer: the refinement delta = refine - sample carries the tolerance (rel-L2 <= 1e-2 of the output,
    # and <= 5e-2 of the delta itself)
    ref = g["refine_pose_repr"]
    o = out["refine_pose_repr"].cpu().numpy()
    x_in = batch["sample_pose_repr"].numpy()
    assert rel_l2(o, ref) <= 1e-2
    assert rel_l2(o - x_in, ref - x_in) <= 5e-2
    # downstream FK of the (slightly different) refined pose
    assert np.abs(out["refine_hand_verts"].cpu().numpy() - g["refine_hand_verts"]).max() < 5e-3
    assert np.abs(out["refine_h2o_dist"].cpu().numpy() - g["refine_h2o_dist"]).max() < 5e-3


def test_r_forward_vs_oracle_unragged_and_single_side():
    """All-right-hand batch (one FK group) at T=24 against the live oracle restatement."""
    from oracle import tamf_oracle as orc
    from tamf_b200 import synth
    m, cfg = _model()
    B, T, P = 3, 24, 512
    batch = synth.make_batch(B, T, nobj=1, seed=9, npoints=P, with_pointcloud=True)
    batch["hand_side"] = ["rh"] * B
    with torch.no_grad():
        ref = orc.r_forward(synth.r_state_dict(cfg, 0), cfg, batch, synth.mano_assets("right"), synth.mano_assets("left"))
    out = m(_dev(batch))
    assert np.abs(out["sample_hand_verts"].cpu().numpy() - ref["sample_hand_verts"].numpy()).max() < 1e-5
    assert np.abs(out["sample_h2o_dist"].cpu().numpy() - ref["sample_h2o_dist"].numpy()).max() < 2e-5
    x_in = batch["sample_pose_repr"].numpy()
    o, r = out["refine_pose_repr"].cpu().numpy(), ref["refine_pose_repr"].numpy()
    assert rel_l2(o, r) <= 1e-2 and rel_l2(o - x_in, r - x_in) <= 5e-2


def test_vertex_normals_vs_oracle():
    import tamf_b200
    from oracle import tamf_oracle as orc
    from tamf_b200 import synth
    A = synth.mano_assets("right")
    rng = np.random.default_rng(4)
    v = (0.05 * rng.standard_normal((5, 778, 3))).astype(np.float32)
    vd, fd = torch.from_numpy(v).cuda(), torch.from_numpy(A["faces"])
    n = tamf_b200.vertex_normals(vd, fd)
    ref = orc.vertex_normals(v, A["faces"])
    # the kernel sums the face contributions exactly (64-bit fixed point) and rounds once; the oracle sums in fp32 in
    # face order: they differ by the oracle's own rounding only
    assert np.abs(n.cpu().numpy() - ref).max() < 2e-5
    for _ in range(20):  # order-independent accumulation: bit-identical from run to run
        assert torch.equal(tamf_b200.vertex_normals(vd, fd), n)
    bad = v.copy()
    bad[1, int(A["faces"][0, 0])] = np.nan
    nb = tamf_b200.vertex_normals(torch.from_numpy(bad).cuda(), fd).cpu().numpy()
    assert np.isnan(nb[1, int(A["faces"][0, 0])]).all() and np.isfinite(nb[0]).all()
    used = np.zeros(778, bool)
    used[np.unique(A["faces"])] = True  # vertices no face references keep a zero normal (eps clamp)
    assert np.abs(np.linalg.norm(n.cpu().numpy(), axis=-1)[:, used] - 1.0).max() < 1e-5


def test_fk_select_matches_full():
    """tamf_mano_fk_select over a subset of frames writes exactly those frames (bitwise equal to the full call)."""
    import tamf_b200
    from tamf_b200 import _lib, synth
    layer = tamf_b200.ManoLayer(side="left", assets=synth.mano_assets("left"))
    rng = np.random.default_rng(2)
    N = 100
    pose = torch.from_numpy(synth.random_pose_repr(rng, 1, N)[0]).cuda()
    betas = torch.from_numpy((0.5 * rng.standard_normal((N, 10))).astype(np.float32)).cuda()
    v_full, j_full = layer.forward_pose_repr(pose, betas)
    ids = torch.tensor([3, 99, 0, 17, 18, 19, 64, 65, 42], dtype=torch.int32, device="cuda")
    v = torch.full((N, 778, 3), -7.0, device="cuda")
    j = torch.full((N, 21, 3), -7.0, device="cuda")
    _lib.check(_lib.lib().tamf_mano_fk_select(layer._handle(pose.device), _lib.POSE_REPR, _lib.ptr(pose), _lib.ptr(betas),
                                              _lib.ptr(ids), ids.numel(), _lib.ptr(v), _lib.ptr(j), _lib.stream_ptr()),
               "fk_select")
    sel = ids.long()
    assert torch.equal(v[sel], v_full[sel]) and torch.equal(j[sel], j_full[sel])
    rest = torch.ones(N, dtype=torch.bool, device="cuda")
    rest[sel] = False
    assert bool((v[rest] == -7.0).all()) and bool((j[rest] == -7.0).all())


def test_r_full_size_vs_oracle():
    """BASELINE.json configs[2] size (arch_refine, B=64, T=160, one object of 8192 points, mixed hand sides) against the
    live oracle: FK of all 10 240 frames <= 1e-5 m, sample_h2o_dist <= 2e-5, refine_pose_repr rel-L2 <= 1e-2 (and the
    refinement delta itself <= 5e-2)."""
    from oracle import tamf_oracle as orc
    from tamf_b200 import synth
    m, cfg = _model()
    B, T = 64, 160
    batch = synth.make_batch(B, T, nobj=1, seed=21, npoints=8192, with_pointcloud=True)
    assert len(set(batch["hand_side"])) == 2
    with torch.no_grad():
        ref = orc.r_forward(synth.r_state_dict(cfg, 0), cfg, batch, synth.mano_assets("right"), synth.mano_assets("left"))
    out = m(_dev(batch))
    for k in ("sample_hand_verts", "sample_hand_joints", "target_hand_verts", "target_hand_joints"):
        e = float((out[k].cpu() - ref[k]).abs().max())
        print(f"{k}: max |d| {e:.2e} m")
        assert e < 1e-5, k
    for k in ("sample_h2o_dist", "target_h2o_dist"):
        e = float((out[k].cpu() - ref[k]).abs().max())
        print(f"{k}: max |d| {e:.2e}")
        assert e < 2e-5, k
    x_in = batch["sample_pose_repr"].numpy()
    o, r = out["refine_pose_repr"].cpu().numpy(), ref["refine_pose_repr"].numpy()
    print(f"refine_pose_repr rel_l2 {rel_l2(o, r):.3e}, delta rel_l2 {rel_l2(o - x_in, r - x_in):.3e}")
    assert rel_l2(o, r) <= 1e-2 and rel_l2(o - x_in, r - x_in) <= 5e-2


def test_r_full_size_properties():
    """BASELINE config 3 size (B=64, T=160, 1 object x 8192 points): finite, batch-row independence."""
    from tamf_b200 import synth
    m, cfg = _model()
    B, T = 64, 160
    batch = synth.make_batch(B, T, nobj=1, seed=1, npoints=8192, with_pointcloud=True)
    out = m(_dev(batch))
    assert all(torch.isfinite(v).all() for v in out.values())
    sub = {k: (v[:2] if isinstance(v, (torch.Tensor, list)) else v) for k, v in batch.items()}
    part = m(_dev(sub))
    assert torch.equal(part["sample_h2o_dist"], out["sample_h2o_dist"][:2])
    assert rel_l2(part["refine_pose_repr"].cpu().numpy(), out["refine_pose_repr"][:2].cpu().numpy()) < 1e-5


def test_r_errors():
    from tamf_b200 import synth
    m, cfg = _model()
    batch = _dev(synth.make_batch(2, 8, nobj=1, seed=0, npoints=64, with_pointcloud=True))
    bad = dict(batch)
    bad["hand_side"] = ["rh", "zz"]
    with pytest.raises(ValueError, match="unexpected hand_side"):
        m(bad)
    with pytest.raises(ValueError, match="unexpected hand_side"):
        m.retrieve_hand_faces("zz")
