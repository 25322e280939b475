"""Drop-in for `chamfer_distance.ChamferDistance` and `point2point_signed`, backed by tamf_nn_query.

Reference: thirdparty/chamfer_distance/chamfer_distance/chamfer_distance.py:65-162 (ChamferDistance.forward ->
(cham_x, cham_y, idx_x, idx_y): squared distances and int64 indices, both directions) and
src/oakink2_tamf/model/loss/chamfer_distance.py:4-64 (point2point_signed)."""
from __future__ import annotations

import torch

from . import _lib


def nn_query(x: torch.Tensor, y: torch.Tensor):
    """x [N,P1,3], y [N,P2,3] CUDA fp32 -> (d2 [N,P1] fp32, idx [N,P1] int64): pytorch3d knn_points(K=1)."""
    if x.ndim != 3 or y.ndim != 3:
        raise ValueError("Expected points to be of shape (N, P, D)")
    if y.shape[0] != x.shape[0] or y.shape[2] != x.shape[2]:
        raise ValueError("y does not have the correct shape.")
    if x.shape[2] != 3:
        raise ValueError("tamf_b200 nearest-neighbour query supports D == 3 only")
    if not x.is_cuda:
        raise RuntimeError("tamf_b200.chamfer needs CUDA tensors (no CPU fallback)")
    x = x.detach().to(torch.float32).contiguous()
    y = y.detach().to(device=x.device, dtype=torch.float32).contiguous()
    N, P1, _ = x.shape
    P2 = y.shape[1]
    d2 = torch.empty((N, P1), dtype=torch.float32, device=x.device)
    idx = torch.empty((N, P1), dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().tamf_nn_query(_lib.ptr(x), _lib.ptr(y), N, P1, P2, _lib.ptr(d2), _lib.ptr(idx),
                                            _lib.stream_ptr(x.device)), "tamf_nn_query")
    return d2, idx


class ChamferDistance(torch.nn.Module):
    """`ChamferDistance()(x, y)` -> (d2_x [N,P1], d2_y [N,P2], idx_x [N,P1], idx_y [N,P2]) (chamfer_distance.py:147-162).
    Only the homogeneous tensor form the reference's call sites use is supported (no lengths / normals / weights)."""

    def forward(self, x, y, x_lengths=None, y_lengths=None, x_normals=None, y_normals=None, weights=None,
                batch_reduction="mean", point_reduction="mean"):
        if batch_reduction is not None and batch_reduction not in ["mean", "sum"]:
            raise ValueError('batch_reduction must be one of ["mean", "sum"] or None')
        if point_reduction not in ["mean", "sum"]:
            raise ValueError('point_reduction must be one of ["mean", "sum"]')
        if x_lengths is not None or y_lengths is not None or weights is not None:
            raise NotImplementedError("heterogeneous clouds / weights are not on the TaMF path")
        d2x, ix = nn_query(x, y)
        d2y, iy = nn_query(y, x)
        return d2x, d2y, ix, iy


def point2point_signed(x, y, x_normals=None, y_normals=None):
    """model/loss/chamfer_distance.py:4-64 -> (y2x_signed [N,P2], x2y_signed [N,P1], yidx_near [N,P2])."""
    N, P1, D = x.shape
    P2 = y.shape[1]
    if y.shape[0] != N or y.shape[2] != D:
        raise ValueError("y does not have the correct shape.")
    _, _, xidx_near, yidx_near = ChamferDistance()(x, y)
    xe = xidx_near.view(N, P1, 1).expand(N, P1, D)
    ye = yidx_near.view(N, P2, 1).expand(N, P2, D)
    x2y = x - y.gather(1, xe)
    y2x = y - x.gather(1, ye)
    if x_normals is not None:
        y_nn = x_normals.gather(1, ye)
        in_out = torch.bmm(y_nn.reshape(-1, 1, 3), y2x.reshape(-1, 3, 1)).reshape(N, -1).sign()
        y2x_signed = y2x.norm(dim=2) * in_out
    else:
        y2x_signed = y2x.norm(dim=2)
    if y_normals is not None:
        x_nn = y_normals.gather(1, xe)
        in_out_x = torch.bmm(x_nn.reshape(-1, 1, 3), x2y.reshape(-1, 3, 1)).reshape(N, -1).sign()
        x2y_signed = x2y.norm(dim=2) * in_out_x
    else:
        x2y_signed = x2y.norm(dim=2)
    return y2x_signed, x2y_signed, yidx_near


class H2OIndex:
    """Canonical object clouds of one batch on the device plus the block index of the pruned exact search
    (tamf_h2o_index_build).  Build once, query with h2o_dist(..., index=...) as often as needed."""

    def __init__(self, obj_points_list, device, build: bool = True):
        import numpy as np
        first = [0]
        for o in obj_points_list:
            first.append(first[-1] + int(o.shape[0]))
        self.first = torch.tensor(first, dtype=torch.int32)
        self.P = int(obj_points_list[0].shape[1])
        self.total_obj = first[-1]
        self.pts = torch.from_numpy(np.concatenate([np.asarray(o, np.float32) for o in obj_points_list], 0)).to(
            device).contiguous()
        self.index = None
        nbytes = _lib.lib().tamf_h2o_index_bytes(self.total_obj, self.P) if build else 0
        if nbytes:
            self.index = torch.empty(nbytes, dtype=torch.uint8, device=device)
            with torch.cuda.device(device):
                _lib.check(_lib.lib().tamf_h2o_index_build(_lib.ptr(self.pts), self.total_obj, self.P,
                                                           _lib.ptr(self.index), nbytes, _lib.stream_ptr(device)),
                           "tamf_h2o_index_build")


def h2o_dist(verts: torch.Tensor, obj_traj: torch.Tensor, obj_points_list, return_idx: bool = False,
             exhaustive: bool = False, index: "H2OIndex | None" = None):
    """Fused `SegmentRefineModel.multi_object_h2o_dist` (segment_refine_model.py:142-168).
    verts [B,T,V,3] CUDA; obj_traj [B,nobj_max,T,9]; obj_points_list: list of B arrays [nobj_b,P,3] -> [B,T,V].
    `index`: a prebuilt H2OIndex of the same clouds (obj_points_list is then ignored).
    `exhaustive=True` runs the brute-force scan (tamf_h2o_dist_exhaustive) instead of the block-pruned exact search --
    same results bit for bit; tests use it as the cross-check."""
    dev = verts.device
    B, T, V, _ = verts.shape
    if index is None:
        index = H2OIndex(obj_points_list, dev, build=not exhaustive)
    verts = verts.detach().to(torch.float32).contiguous()
    obj_traj = obj_traj.detach().to(device=dev, dtype=torch.float32).contiguous()
    dist = torch.empty((B, T, V), dtype=torch.float32, device=dev)
    idx = torch.empty((B, T, V), dtype=torch.int64, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        if exhaustive or index.index is None:
            fn, third = (L.tamf_h2o_dist_exhaustive if exhaustive else L.tamf_h2o_dist), index.pts
        else:
            fn, third = L.tamf_h2o_dist_indexed, index.index
        _lib.check(fn(_lib.ptr(verts), _lib.ptr(obj_traj), _lib.ptr(third), _lib.C.c_void_p(index.first.data_ptr()), B, T,
                      V, obj_traj.shape[1], index.P, _lib.ptr(dist), _lib.ptr(idx), _lib.stream_ptr(dev)),
                   "tamf_h2o_dist")
    return (dist, idx) if return_idx else dist
