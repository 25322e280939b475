"""Drop-in for manotorch `ManoLayer` in the mode the reference uses (quat, center_idx=0, no PCA, flat mean).

Reference: thirdparty/manotorch/manotorch/manolayer.py:13-25 (MANOOutput), :39-98 (__init__ / buffers),
:268-285 (forward), :329-359 (get_mano_closed_faces)."""
from __future__ import annotations

import ctypes as C
import os
from collections import namedtuple
from typing import Optional

import numpy as np
import torch

from . import _lib

MANOOutput = namedtuple(
    "MANOOutput", ["verts", "joints", "center_idx", "center_joint", "full_poses", "betas", "transforms_abs"])
MANOOutput.__new__.__defaults__ = (None,) * len(MANOOutput._fields)


def _load_assets(mano_assets_root: str, side: str) -> dict:
    """MANO_{SIDE}.pkl needs chumpy to unpickle (absent / broken on py3.12); a .npz with the same arrays
    (shapedirs, posedirs, v_template, J_regressor, weights, f) next to it is accepted instead."""
    base = os.path.join(mano_assets_root, "models", f"MANO_{side.upper()}")
    if os.path.isfile(base + ".npz"):
        z = np.load(base + ".npz")
        return dict(shapedirs=z["shapedirs"], posedirs=z["posedirs"], v_template=z["v_template"],
                    J_regressor=z["J_regressor"], weights=z["weights"], faces=z["f"])
    assert os.path.isfile(base + ".pkl"), f"Can not find MANO assets {base}.pkl, please follow steps in README.md"
    import pickle
    dd = pickle.load(open(base + ".pkl", "rb"), encoding="latin1")
    J = dd["J_regressor"]
    J = J.toarray() if hasattr(J, "toarray") else np.asarray(J)
    g = lambda k: np.asarray(getattr(dd[k], "r", dd[k]))
    return dict(shapedirs=g("shapedirs"), posedirs=g("posedirs"), v_template=g("v_template"), J_regressor=J,
                weights=g("weights"), faces=np.asarray(dd["f"]))


def quaternion_to_axis_angle(quaternions: torch.Tensor) -> torch.Tensor:
    """(w,x,y,z) quaternions [...,4] -> axis-angle [...,3] (manotorch/utils/geometry.py:276-300; host-side glue of
    `MANOOutput.full_poses`, elementwise torch ops)."""
    norms = torch.norm(quaternions[..., 1:], p=2, dim=-1, keepdim=True)
    half = torch.atan2(norms, quaternions[..., :1])
    ang = 2 * half
    small = ang.abs() < 1e-6
    safe = torch.where(small, torch.ones_like(ang), ang)
    ratio = torch.where(small, 0.5 - (ang * ang) / 48, torch.sin(half) / safe)
    return quaternions[..., 1:] / ratio


class ManoLayer(torch.nn.Module):
    def __init__(self, rot_mode: str = "quat", side: str = "right", center_idx: Optional[int] = 0,
                 mano_assets_root: str = "assets/mano", use_pca: bool = False, flat_hand_mean: bool = True,
                 ncomps: int = 15, assets: Optional[dict] = None, **kargs):
        super().__init__()
        if rot_mode != "quat":
            raise NotImplementedError(f"tamf_b200.ManoLayer implements rot_mode='quat' (the TaMF path), got {rot_mode}")
        if center_idx != 0 or use_pca or not flat_hand_mean:
            raise NotImplementedError("tamf_b200.ManoLayer implements center_idx=0, use_pca=False, flat_hand_mean=True")
        self.center_idx, self.rot_mode, self.side, self.use_pca = center_idx, rot_mode, side, use_pca
        self.mano_assets_root, self.flat_hand_mean, self.ncomps, self.rot_dim = mano_assets_root, flat_hand_mean, -1, 4
        A = assets if assets is not None else _load_assets(mano_assets_root, side)
        f32 = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=np.float32))
        self.register_buffer("th_betas", torch.zeros(1, 10))
        self.register_buffer("th_shapedirs", f32(A["shapedirs"]))
        self.register_buffer("th_posedirs", f32(A["posedirs"]))
        self.register_buffer("th_v_template", f32(A["v_template"]).unsqueeze(0))
        self.register_buffer("th_J_regressor", f32(A["J_regressor"]))
        self.register_buffer("th_weights", f32(A["weights"]))
        self.register_buffer("th_faces", torch.from_numpy(np.asarray(A["faces"]).astype(np.int64)))
        self._handles = {}  # device index -> tamf_mano*

    def _handle(self, device: torch.device):
        key = device.index if device.index is not None else torch.cuda.current_device()
        if key not in self._handles:
            h = C.c_void_p()
            cpu = lambda t: t.detach().cpu().contiguous()
            sd, pd, vt = cpu(self.th_shapedirs), cpu(self.th_posedirs), cpu(self.th_v_template[0])
            jr, w = cpu(self.th_J_regressor), cpu(self.th_weights)
            with torch.cuda.device(key):
                _lib.check(_lib.lib().tamf_mano_create(_lib.ptr(sd), _lib.ptr(pd), _lib.ptr(vt), _lib.ptr(jr),
                                                       _lib.ptr(w), 1 if self.side == "right" else 0, C.byref(h)),
                           "tamf_mano_create")
            self._handles[key] = h
        return self._handles[key]

    def __del__(self):
        try:
            for h in self._handles.values():
                _lib.lib().tamf_mano_destroy(h)
        except Exception:
            pass

    def _fk(self, mode: int, pose: torch.Tensor, betas: torch.Tensor, full: bool = False):
        if not pose.is_cuda:
            raise RuntimeError("tamf_b200.ManoLayer needs CUDA tensors (no CPU fallback)")
        dev = pose.device
        N = pose.shape[0]
        pose = pose.detach().to(torch.float32).contiguous()
        betas = betas.detach().to(device=dev, dtype=torch.float32).contiguous()
        if betas.shape != (N, 10):
            raise ValueError(f"betas must be [N,10], got {tuple(betas.shape)}")
        verts = torch.empty((N, 778, 3), dtype=torch.float32, device=dev)
        joints = torch.empty((N, 21, 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            if full:
                center = torch.empty((N, 1, 3), dtype=torch.float32, device=dev)
                transf = torch.empty((N, 16, 4, 4), dtype=torch.float32, device=dev)
                _lib.check(_lib.lib().tamf_mano_fk_full(self._handle(dev), mode, _lib.ptr(pose), _lib.ptr(betas), N,
                                                        _lib.ptr(verts), _lib.ptr(joints), _lib.ptr(center),
                                                        _lib.ptr(transf), _lib.stream_ptr(dev)), "tamf_mano_fk_full")
                return verts, joints, center, transf
            _lib.check(_lib.lib().tamf_mano_fk(self._handle(dev), mode, _lib.ptr(pose), _lib.ptr(betas), N,
                                               _lib.ptr(verts), _lib.ptr(joints), _lib.stream_ptr(dev)), "tamf_mano_fk")
        return verts, joints

    def forward(self, pose_coeffs: torch.Tensor, betas: Optional[torch.Tensor] = None, **kwargs):
        """pose_coeffs [N,16,4] (or [N,64]) quaternions (w,x,y,z); betas [N,10] -> MANOOutput (manolayer.py:268-285):
        verts / joints root-centred, `center_joint` [N,1,3] the root joint before centring, `transforms_abs`
        [N,16,4,4] the centre-shifted global joint transforms (:251-258), `full_poses` [N,48] axis-angle (:119-126)."""
        N = pose_coeffs.shape[0]
        if pose_coeffs.numel() != N * 64:
            raise ValueError(f"pose_coeffs must be [N,16,4], got {tuple(pose_coeffs.shape)}")
        if betas is None:
            betas = self.th_betas.to(pose_coeffs.device).expand(N, 10)
        verts, joints, center, transf = self._fk(_lib.POSE_QUAT, pose_coeffs.reshape(N, 64), betas, full=True)
        full_poses = quaternion_to_axis_angle(pose_coeffs.reshape(N, 16, 4).to(torch.float32)).reshape(N, -1)
        return MANOOutput(verts=verts, joints=joints, center_idx=self.center_idx, center_joint=center,
                          full_poses=full_poses, betas=betas, transforms_abs=transf)

    def forward_pose_repr(self, pose_repr: torch.Tensor, betas: torch.Tensor):
        """pose_repr [N,99] (tsl + 16 x rot6d) -> world verts [N,778,3], joints [N,21,3]: the per-item body of
        SegmentRefineModel.batch_recover_mano_from_pose_repr (segment_refine_model.py:117-131) in one kernel."""
        return self._fk(_lib.POSE_REPR, pose_repr, betas)

    def get_mano_closed_faces(self):
        """manolayer.py:329-359 (wrist-closing faces appended)."""
        close_faces = torch.tensor([
            [92, 38, 122], [234, 92, 122], [239, 234, 122], [279, 239, 122], [215, 279, 122], [215, 122, 118],
            [215, 118, 117], [215, 117, 119], [215, 119, 120], [215, 120, 108], [215, 108, 79], [215, 79, 78],
            [215, 78, 121], [214, 215, 121]], dtype=torch.long)
        return torch.cat([self.th_faces.clone().detach().cpu(), close_faces])
