"""Drop-in for `SegmentRefineModel` (MF-MDM R): same constructor arguments, state_dict keys, `forward(batch)` and
13-key result dict as src/oakink2_tamf/model/segment_refine_model.py:21-250, running on libtamf_b200.

One forward = 3 x (MANO FK + vertex normals + fused hand->object distance) + one transformer pass
(segment_refine_model.py:193-201, :207-217, :220-232).  The reference loops over batch items and objects in Python;
here every stage is one batched kernel over all B*T frames (right and left hands are selected by frame index, the
object transform is fused into the nearest-neighbour scan)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .chamfer import H2OIndex, h2o_dist
from .manolayer import ManoLayer
from .mdm import InterationSegmentMDM, _Holder, _positional_table, build_encoder_container, layer_weight_structs


def vertex_normals(verts: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    """verts [N,V,3] CUDA fp32, faces [F,3] -> area-weighted unit vertex normals [N,V,3] (tamf_vertex_normals)."""
    if not verts.is_cuda:
        raise RuntimeError("tamf_b200.vertex_normals needs CUDA tensors (no CPU fallback)")
    verts = verts.detach().to(torch.float32).contiguous()
    f32 = faces.to(device=verts.device, dtype=torch.int32).contiguous()
    N, V, _ = verts.shape
    out = torch.empty_like(verts)
    with torch.cuda.device(verts.device):
        _lib.check(_lib.lib().tamf_vertex_normals(_lib.ptr(verts), _lib.ptr(f32), N, V, f32.shape[0], _lib.ptr(out),
                                                  _lib.stream_ptr(verts.device)), "tamf_vertex_normals")
    return out


class SegmentRefineModel(nn.Module):
    def __init__(self, mano_path, input_dim=99, obj_input_dim=9, hand_shape_dim=10, obj_embed_dim=768, latent_dim=256,
                 ff_size=1024, num_layers=8, num_heads=4, dropout=0.1, activation="gelu", use_pc=False,
                 mano_assets: Optional[dict] = None):
        """`mano_assets` = {"right": {...}, "left": {...}} bypasses the MANO .pkl files (synthetic assets offline)."""
        super().__init__()
        if activation != "gelu":
            raise NotImplementedError("tamf_b200 implements activation='gelu' (config/arch_refine.yml)")
        self.latent_dim, self.ff_size, self.num_layers, self.num_heads = latent_dim, ff_size, num_layers, num_heads
        self.dropout, self.activation = dropout, activation
        self.input_feats, self.obj_input_feats = input_dim, obj_input_dim
        self.hand_shape_feats, self.obj_embed_feats = hand_shape_dim, obj_embed_dim
        self.use_pc = use_pc
        ma = mano_assets or {}
        mk = lambda side: ManoLayer(mano_assets_root=mano_path, rot_mode="quat", side=side, center_idx=0, use_pca=False,
                                    flat_hand_mean=True, assets=ma.get(side))
        self.mano_layer_rh, self.mano_layer_lh = mk("right"), mk("left")
        d = latent_dim
        self.hand_side_process = _Holder()
        self.hand_side_process.register_buffer("rh_embed", torch.zeros(d))
        lh = torch.zeros(d)
        lh[0] = 1.0
        self.hand_side_process.register_buffer("lh_embed", lh)
        self.hand_shape_process = _Holder()
        self.hand_shape_process.shape_embed = nn.Linear(hand_shape_dim, d)
        self.obj_embed_process = _Holder()
        self.obj_embed_process.embedding = nn.Linear(obj_embed_dim, d)
        self.input_process = _Holder()
        self.input_process.poseEmbedding = nn.Linear(input_dim, d)
        self.obj_input_process = _Holder()
        self.obj_input_process.poseEmbedding = nn.Linear(obj_input_dim, d)
        self.h2o_dist_input_process = _Holder()
        self.h2o_dist_input_process.poseEmbedding = nn.Linear(778, d)
        self.input_merge = nn.Sequential(nn.Linear(d * 3, d), nn.SiLU(), nn.Linear(d, d))
        self.sequence_pos_encoder = _Holder()
        self.sequence_pos_encoder.register_buffer("pe", _positional_table(d))
        self.seqTransEncoder = build_encoder_container(d, num_heads, ff_size, dropout, activation, num_layers)
        self.output_process = _Holder()
        self.output_process.poseFinal = nn.Linear(d, input_dim)
        self._handle, self._handle_dev, self._ws, self._bound = None, None, None, None

    # ---- reference surface ----
    def retrieve_hand_faces(self, hand_side):
        if hand_side == "rh":
            return self.mano_layer_rh.th_faces
        elif hand_side == "lh":
            return self.mano_layer_lh.th_faces
        raise ValueError(f"unexpected hand_side: {hand_side}")

    def load_state_dict(self, state_dict, strict=True, **kw):
        res = super().load_state_dict(state_dict, strict=strict, **kw)
        self._drop_handle()
        return res

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._drop_handle()
        return r

    def _drop_handle(self):
        if getattr(self, "_handle", None) is not None:
            _lib.lib().tamf_refiner_destroy(self._handle)
        self._handle, self._bound, self._ws = None, None, None

    def __del__(self):
        try:
            self._drop_handle()
        except Exception:
            pass

    # ---- library plumbing ----
    def _ensure_handle(self, device):
        if device.type != "cuda":
            raise RuntimeError("tamf_b200.SegmentRefineModel runs on a B200 (model.to('cuda')); no CPU fallback")
        if self._handle is not None and self._handle_dev == device:
            return
        self._drop_handle()
        keep = []
        host = lambda t: keep.append(t.detach().to("cpu", torch.float32).contiguous()) or keep[-1].data_ptr()
        cfg = _lib.TamfCfg(self.input_feats, self.obj_input_feats, self.hand_shape_feats, self.obj_embed_feats,
                           self.latent_dim, self.ff_size, self.num_layers, self.num_heads, 0, 0)
        w = _lib.TamfRWeights()
        pairs = dict(shape=self.hand_shape_process.shape_embed, objemb=self.obj_embed_process.embedding,
                     pose=self.input_process.poseEmbedding, objtraj=self.obj_input_process.poseEmbedding,
                     dist=self.h2o_dist_input_process.poseEmbedding, merge0=self.input_merge[0],
                     merge2=self.input_merge[2], final=self.output_process.poseFinal)
        for k, lin in pairs.items():
            setattr(w, k + "_w", host(lin.weight))
            setattr(w, k + "_b", host(lin.bias))
        pe = self.sequence_pos_encoder.pe[:, 0]
        w.pe, w.pe_rows = host(pe), pe.shape[0]
        layers = layer_weight_structs(self.seqTransEncoder, keep)
        w.layers = C.cast(layers, C.POINTER(_lib.TamfLayerWeights))
        h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(_lib.lib().tamf_refiner_create(C.byref(cfg), C.byref(w), C.byref(h)), "tamf_refiner_create")
        self._handle, self._handle_dev = h, device

    def _ensure_bound(self, B, T, device):
        self._ensure_handle(device)
        if self._bound == (B, T):
            return
        L = _lib.lib()
        nbytes = L.tamf_refiner_workspace_bytes(self._handle, B, T)
        self._ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        base = (self._ws.data_ptr() + 255) & ~255
        with torch.cuda.device(device):
            _lib.check(L.tamf_refiner_bind(self._handle, B, T, C.c_void_p(base), nbytes), "tamf_refiner_bind")
        self._bound = (B, T)

    def batch_recover_mano_from_pose_repr(self, batch_pose_repr, batch_shape, batch_hand_side):
        """segment_refine_model.py:107-140 -> (verts [B,T,778,3], joints [B,T,21,3], normals [B,T,778,3]); one FK launch
        per hand side over the frames of that side (tamf_mano_fk_select), one normals launch per side."""
        dev = batch_pose_repr.device
        B, T, _ = batch_pose_repr.shape
        side_ids = InterationSegmentMDM.hand_side_ids(batch_hand_side)  # raises ValueError like :129
        pose = batch_pose_repr.detach().to(torch.float32).contiguous().view(B * T, -1)
        betas = batch_shape.detach().to(device=dev, dtype=torch.float32).contiguous().view(B * T, 10)
        verts = torch.empty((B * T, 778, 3), dtype=torch.float32, device=dev)
        joints = torch.empty((B * T, 21, 3), dtype=torch.float32, device=dev)
        normals = torch.empty((B * T, 778, 3), dtype=torch.float32, device=dev)
        frame = torch.arange(B * T, dtype=torch.int32).view(B, T)
        for sid, layer in ((0, self.mano_layer_rh), (1, self.mano_layer_lh)):
            rows = [b for b in range(B) if side_ids[b] == sid]
            if not rows:
                continue
            ids = frame[rows].reshape(-1).to(dev)
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().tamf_mano_fk_select(layer._handle(dev), _lib.POSE_REPR, _lib.ptr(pose),
                                                          _lib.ptr(betas), _lib.ptr(ids), ids.numel(), _lib.ptr(verts),
                                                          _lib.ptr(joints), _lib.stream_ptr(dev)), "tamf_mano_fk_select")
            if len(rows) == B:
                normals = vertex_normals(verts, layer.th_faces)
            else:
                idl = ids.long()
                normals[idl] = vertex_normals(verts[idl], layer.th_faces)
        return verts.view(B, T, 778, 3), joints.view(B, T, 21, 3), normals.view(B, T, 778, 3)

    def multi_object_h2o_dist(self, batch_hand_verts, batch_hand_normals, batch_obj_list, batch_obj_traj,
                              batch_obj_verts_list, index=None):
        """segment_refine_model.py:142-168 -> [B,T,778] unsigned hand->object distance (normals only feed the
        discarded y2x_signed, :165).  `index`: an H2OIndex of the same clouds built by the caller (forward() builds
        one for its three queries)."""
        if index is None:
            index = self.object_index(batch_obj_list, batch_obj_verts_list, batch_hand_verts.device)
        return h2o_dist(batch_hand_verts, batch_obj_traj, None, index=index)

    @staticmethod
    def object_index(batch_obj_list, batch_obj_verts_list, device):
        pts = [np.asarray(o, np.float32)[: len(l)] for o, l in zip(batch_obj_verts_list, batch_obj_list)]
        return H2OIndex(pts, device)

    def forward(self, batch):
        x_in = batch["sample_pose_repr"]
        if not x_in.is_cuda:
            raise RuntimeError("tamf_b200.SegmentRefineModel needs CUDA tensors (no CPU fallback)")
        dev = x_in.device
        B, T, nf = x_in.shape
        if nf != self.input_feats:
            raise ValueError(f"sample_pose_repr must be [B,T,{self.input_feats}], got {tuple(x_in.shape)}")
        self._ensure_bound(B, T, dev)
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        x_in, shape, traj, emb = f32(x_in), f32(batch["shape"]), f32(batch["obj_traj"]), f32(batch["obj_embedding"])
        obj_pts = batch["obj_pointcloud"] if self.use_pc else batch["obj_verts"]
        side = torch.tensor(InterationSegmentMDM.hand_side_ids(batch["hand_side"]), dtype=torch.int32, device=dev)
        hv, hj, hn = self.batch_recover_mano_from_pose_repr(x_in, shape, batch["hand_side"])
        oix = self.object_index(batch["obj_list"], obj_pts, dev)  # clouds uploaded + indexed once for the three queries
        h2o = self.multi_object_h2o_dist(hv, hn, batch["obj_list"], traj, obj_pts, index=oix)
        out = torch.empty_like(x_in)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().tamf_refiner_forward(self._handle, _lib.ptr(x_in), _lib.ptr(h2o), _lib.ptr(side),
                                                       _lib.ptr(shape), _lib.ptr(traj), _lib.ptr(emb), traj.shape[1],
                                                       _lib.ptr(out), _lib.stream_ptr(dev)), "tamf_refiner_forward")
        rv, rj, rn = self.batch_recover_mano_from_pose_repr(out, shape, batch["hand_side"])
        r_h2o = self.multi_object_h2o_dist(rv, rn, batch["obj_list"], traj, obj_pts, index=oix)
        tv, tj, tn = self.batch_recover_mano_from_pose_repr(f32(batch["pose_repr"]), shape, batch["hand_side"])
        t_h2o = self.multi_object_h2o_dist(tv, tn, batch["obj_list"], traj, obj_pts, index=oix)
        return {
            "refine_pose_repr": out, "refine_hand_verts": rv, "refine_hand_joints": rj, "refine_hand_normals": rn,
            "refine_h2o_dist": r_h2o, "target_hand_verts": tv, "target_hand_joints": tj, "target_hand_normals": tn,
            "target_h2o_dist": t_h2o, "sample_hand_verts": hv, "sample_hand_joints": hj, "sample_hand_normals": hn,
            "sample_h2o_dist": h2o,
        }
