"""Chamfer NN parity (bit-exact indices and squared distances) against the oracle -- through the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(x, y):
    import tamf_b200
    d2, idx = tamf_b200.nn_query(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda())
    return d2.cpu().numpy(), idx.cpu().numpy()


@pytest.mark.parametrize("N,P1,P2", [(1, 1, 1), (2, 778, 8192), (3, 7, 1000), (5, 778, 16384), (1, 4096, 37),
                                     (64, 100, 1025), (2, 4097, 300), (3, 16384, 778), (1, 10001, 1500)])
def test_nn_bit_exact(N, P1, P2):
    from oracle import tamf_oracle as orc
    rng = np.random.default_rng(N * 1000 + P1 + P2)
    x = (0.1 * rng.standard_normal((N, P1, 3))).astype(np.float32)
    y = (0.05 * rng.standard_normal((N, P2, 3))).astype(np.float32)
    d2, idx = _run(x, y)
    rd2, ridx = orc.nn_query(x, y)
    assert idx.dtype == np.int64 and d2.dtype == np.float32
    assert np.array_equal(idx, ridx)
    assert np.array_equal(d2.view(np.uint32), rd2.view(np.uint32))  # bit-exact squared distances


def test_nn_ties_lowest_index():
    """Duplicated candidates and a lattice (many exact ties): the lowest index must win (documented rule)."""
    from oracle import tamf_oracle as orc
    rng = np.random.default_rng(0)
    base = rng.integers(-3, 4, (1, 300, 3)).astype(np.float32) * 0.25
    y = np.concatenate([base, base, base], 1)           # every candidate appears three times
    x = rng.integers(-3, 4, (1, 500, 3)).astype(np.float32) * 0.25
    d2, idx = _run(x, y)
    rd2, ridx = orc.nn_query(x, y)
    assert np.array_equal(idx, ridx)
    assert idx.max() < 300                               # never one of the later duplicates
    assert np.array_equal(d2, rd2)


def test_nn_numpy_and_c_oracle_agree():
    from oracle import tamf_oracle as orc
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 50, 3)).astype(np.float32)
    y = rng.standard_normal((2, 400, 3)).astype(np.float32)
    a, b = orc.nn_query(x, y), orc.nn_query_numpy(x, y)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_nn_empty_and_errors():
    import tamf_b200
    x = torch.zeros((2, 0, 3), device="cuda")
    y = torch.zeros((2, 5, 3), device="cuda")
    d2, idx = tamf_b200.nn_query(x, y)
    assert d2.shape == (2, 0) and idx.shape == (2, 0)
    with pytest.raises(ValueError):
        tamf_b200.nn_query(torch.zeros((2, 4, 3), device="cuda"), torch.zeros((2, 0, 3), device="cuda"))
    with pytest.raises(ValueError):
        tamf_b200.nn_query(torch.zeros((2, 4, 3), device="cuda"), torch.zeros((3, 4, 3), device="cuda"))


def test_chamfer_distance_api_and_point2point():
    """ChamferDistance()(x, y) 4-tuple and point2point_signed against the oracle's statement of
    model/loss/chamfer_distance.py:36-62."""
    import tamf_b200
    from oracle import tamf_oracle as orc
    rng = np.random.default_rng(9)
    x = (0.1 * rng.standard_normal((4, 778, 3))).astype(np.float32)
    y = (0.1 * rng.standard_normal((4, 2048, 3))).astype(np.float32)
    xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    dx, dy, ix, iy = tamf_b200.ChamferDistance()(xt, yt)
    rdx, rix = orc.nn_query(x, y)
    rdy, riy = orc.nn_query(y, x)
    assert np.array_equal(ix.cpu().numpy(), rix) and np.array_equal(iy.cpu().numpy(), riy)
    assert np.array_equal(dx.cpu().numpy(), rdx) and np.array_equal(dy.cpu().numpy(), rdy)
    y2x, x2y, yidx = tamf_b200.point2point_signed(xt, yt)
    ref_x2y, _ = orc.point2point_h2o(torch.from_numpy(x), torch.from_numpy(y))
    assert np.array_equal(yidx.cpu().numpy(), riy)
    np.testing.assert_allclose(x2y.cpu().numpy(), ref_x2y.numpy(), rtol=0, atol=1e-7)
    # size-independent property: the reported distance is attained by the reported index, and no candidate is closer
    d_all = torch.cdist(xt[:1], yt[:1])[0]
    assert torch.all(d_all.min(dim=1).values + 1e-6 >= x2y[0])


def test_chamfer_reference_call_shape_both_directions(golden):
    """`ChamferDistance()(x[T,778,3], y[T,2*8192,3])` and `point2point_signed(..., x_normals, y_normals)` at the
    reference's own call shape (segment_refine_model.py:165, interaction_segment_extra_loss.py:157): the reverse pass
    queries with the 16 384 object points.  Checked against (a) outputs of the REFERENCE's code
    (tests/golden/p2p_signed.npz: chamfer_distance.py:147-162 + model/loss/chamfer_distance.py:4-64 over the knn stub)
    and (b) the live oracle -- indices and squared distances bit-exact."""
    import tamf_b200
    from oracle import tamf_oracle as orc
    from tamf_b200 import synth
    g = golden("p2p_signed.npz")
    x, xn, y, yn = synth.p2p_clouds(seed=int(g["seed"]), T=int(g["T"]), nobj=int(g["nobj"]), P=int(g["P"]))
    xt, xnt, yt, ynt = (torch.from_numpy(a).cuda() for a in (x, xn, y, yn))
    cx, cy, ix, iy = tamf_b200.ChamferDistance()(xt, yt)
    assert ix.dtype == torch.int64 and iy.shape == (x.shape[0], y.shape[1])
    assert np.array_equal(ix.cpu().numpy(), g["idx_x"]) and np.array_equal(iy.cpu().numpy(), g["idx_y"])
    assert np.array_equal(cx.cpu().numpy(), g["cham_x"]) and np.array_equal(cy.cpu().numpy(), g["cham_y"])
    rdy, riy = orc.nn_query(y, x)
    assert np.array_equal(iy.cpu().numpy(), riy) and np.array_equal(cy.cpu().numpy(), rdy)
    y2x, x2y, yidx = tamf_b200.point2point_signed(xt, yt, x_normals=xnt, y_normals=ynt)
    assert np.array_equal(yidx.cpu().numpy(), g["yidx_near"])
    # |d| to 1e-7; the sign is that of a 3-term dot product: compare it wherever the product is not within rounding of 0
    for ours, ref in ((y2x, g["y2x_signed"]), (x2y, g["x2y_signed"])):
        o = ours.cpu().numpy()
        np.testing.assert_allclose(np.abs(o), np.abs(ref), rtol=0, atol=1e-7)
        clear = np.abs(ref) > 1e-6
        assert np.array_equal(np.sign(o[clear]), np.sign(ref[clear]))
    assert (g["y2x_signed"] < 0).any() and (g["y2x_signed"] > 0).any()  # the signed branch is exercised
    y2x_u, x2y_u, _ = tamf_b200.point2point_signed(xt, yt)
    np.testing.assert_allclose(y2x_u.cpu().numpy(), g["y2x_unsigned"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(x2y_u.cpu().numpy(), g["x2y_unsigned"], rtol=0, atol=1e-7)
    # the oracle's full restatement agrees with the reference's code as well
    r_y2x, r_x2y, r_idx = orc.point2point_signed(*(torch.from_numpy(a) for a in (x, y, xn, yn)))
    assert np.array_equal(r_idx.numpy(), g["yidx_near"])
    np.testing.assert_allclose(r_y2x.numpy(), g["y2x_signed"], rtol=0, atol=1e-7)


def test_h2o_fused_matches_materialised():
    """tamf_h2o_dist (transform fused into the scan) vs oracle on materialised world points: distances by value
    (1e-6 m), indices equal wherever the two nearest candidates are not within rounding of each other."""
    import tamf_b200
    from oracle import tamf_oracle as orc
    from tamf_b200 import synth
    B, T, P = 3, 5, 1024
    batch = synth.make_batch(B, T, nobj=2, seed=4, ragged=True, npoints=P, with_pointcloud=True)
    rng = np.random.default_rng(5)
    verts = torch.from_numpy((0.1 * rng.standard_normal((B, T, 778, 3))).astype(np.float32))
    obj_num = [len(o) for o in batch["obj_list"]]
    ref = orc.h2o_dist(verts, batch["obj_traj"], obj_num, batch["obj_pointcloud"])
    out = tamf_b200.h2o_dist(verts.cuda(), batch["obj_traj"].cuda(), batch["obj_pointcloud"])
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=0, atol=1e-6)


def _h2o_both(verts, traj, clouds):
    import tamf_b200
    a = tamf_b200.h2o_dist(verts.cuda(), traj.cuda(), clouds, return_idx=True)
    b = tamf_b200.h2o_dist(verts.cuda(), traj.cuda(), clouds, return_idx=True, exhaustive=True)
    return [t.cpu().numpy() for t in a], [t.cpu().numpy() for t in b]


def _assert_same(a, b):
    assert np.array_equal(a[1], b[1]), f"{(a[1] != b[1]).sum()} indices differ"
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))  # bit-exact distances (NaN patterns included)


@pytest.mark.parametrize("B,T,P,nobj,scale", [(3, 5, 1024, 2, 0.1), (2, 3, 8192, 1, 0.1), (2, 4, 1000, 3, 0.02),
                                              (1, 2, 37, 1, 0.3), (2, 2, 64, 2, 1.0), (4, 6, 4096, 2, 0.005)])
def test_h2o_pruned_equals_exhaustive(B, T, P, nobj, scale):
    """The block-pruned exact search must reproduce the exhaustive scan bit for bit -- far hands, hands inside the
    cloud (scale << cloud size) and ragged object counts."""
    from tamf_b200 import synth
    batch = synth.make_batch(B, T, nobj=nobj, seed=P + B, ragged=nobj > 1, npoints=P, with_pointcloud=True)
    rng = np.random.default_rng(P * 7 + T)
    verts = torch.from_numpy((scale * rng.standard_normal((B, T, 778, 3))).astype(np.float32))
    # half of the vertices sit next to object points (contact): exercises small nearest distances
    for b in range(B):
        pc = np.asarray(batch["obj_pointcloud"][b], np.float32)[0]
        tr = batch["obj_traj"][b, 0].numpy()
        from oracle import tamf_oracle as orc
        R = orc.rot6d_to_rotmat(torch.from_numpy(tr[:, 3:9])).numpy()  # rows b1,b2,b3 (transform.py:148-154)
        for t in range(T):
            sel = rng.integers(0, pc.shape[0], 389)
            world = pc[sel] @ R[t].T + tr[t, :3]
            verts[b, t, :389] = torch.from_numpy(world + 1e-3 * rng.standard_normal((389, 3)).astype(np.float32))
    a, bb = _h2o_both(verts, batch["obj_traj"], batch["obj_pointcloud"])
    _assert_same(a, bb)


def test_h2o_pruned_ties_degenerate_and_nan():
    """Exact ties (lattice cloud with every point duplicated, axis-aligned transforms), a degenerate rot6d (pruning
    must switch itself off), NaN vertices."""
    rng = np.random.default_rng(1)
    B, T, P = 2, 3, 512
    base = rng.integers(-4, 5, (1, P // 2, 3)).astype(np.float32) * 0.125
    cloud = np.concatenate([base, base], 1)  # every point twice: the lower index must win
    clouds = [cloud.copy(), cloud.copy()]
    traj = torch.zeros(B, 1, T, 9)
    traj[..., 3] = 1.0
    traj[..., 7] = 1.0  # identity rotation
    traj[..., 0:3] = torch.from_numpy(rng.integers(-2, 3, (B, 1, T, 3)).astype(np.float32) * 0.125)
    verts = torch.from_numpy(rng.integers(-6, 7, (B, T, 778, 3)).astype(np.float32) * 0.125)
    a, b = _h2o_both(verts, traj, clouds)
    _assert_same(a, b)
    assert a[1].max() < P // 2
    traj2 = traj.clone()
    traj2[0, 0, 1, 3:9] = 0.0  # degenerate rot6d: R is not a rotation
    traj2[1, 0, 0, 3:9] = torch.tensor([1.0, 0.0, 0.0, 2.0, 0.0, 0.0])  # colinear columns
    a, b = _h2o_both(verts + 0.01 * torch.randn_like(verts), traj2, clouds)
    _assert_same(a, b)
    v3 = verts.clone()
    v3[0, 0, 5, 1] = float("nan")
    v3[1, 2, 700] = float("inf")
    a, b = _h2o_both(v3, traj, clouds)
    _assert_same(a, b)
