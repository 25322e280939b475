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
                                     (64, 100, 1025)])
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
