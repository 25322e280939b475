"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the TaMF sampling hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this module, and only as the checker (never as the thing measured as "ours", never shipped).
The product path (`oakink2-tamf_b200/`) never imports it and fails loudly without its CUDA library.

Every function cites the reference file:line it restates (paths relative to /root/reference).
Arithmetic: fp32 torch-CPU / numpy, i.e. exactly what the reference's PyTorch code executes on CPU.

Pinning status (checked by tests/test_oracle_pin.py and by oracle/make_golden.py in the authoring
container, against the *reference's own modules* imported via oracle/ref_shims.py; the outputs are
committed under tests/golden/ so the pin travels to the GPU box):
  * diffusion schedule / p_sample         : pinned (reference runs unmodified)
  * G denoiser forward (InterationSegmentMDM): pinned (reference module, random-init CLIP stub)
  * rot6d / quat / SE3 helpers             : pinned (reference runs unmodified)
  * ManoLayer FK (quat mode)               : pinned on synthetic MANO-shaped assets (real MANO .pkl is
                                             licensed/absent; `skinning_layer` itself runs verbatim)
  * R forward (SegmentRefineModel)         : pinned modulo the pytorch3d stub below
  * chamfer NN (`knn_points`, K=1) and `Meshes.verts_normals_packed`: **PARITY UNPINNED** --
    pytorch3d==0.7.2 (requirements.dist.txt:331) is neither vendored nor installed; `nn_query` /
    `vertex_normals` below restate its published algorithm and *define* the arithmetic
    (fp32, per-axis diff, (dx*dx + dy*dy) + dz*dz with every op rounded, lowest index on ties).
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# Diffusion schedule + ancestral sampler
# --------------------------------------------------------------------------------------------


def cosine_betas(num_steps: int = 1000, max_beta: float = 0.999) -> np.ndarray:
    """model/diffusion/gaussian_diffusion.py:36-40,45-62 (`get_named_beta_schedule("cosine")`)."""
    ab = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    return np.array(
        [min(1 - ab((i + 1) / num_steps) / ab(i / num_steps), max_beta) for i in range(num_steps)], dtype=np.float64)


def diffusion_tables(num_steps: int = 1000) -> dict:
    """gaussian_diffusion.py:116-161 (`GaussianDiffusion.__init__`), float64; SpacedDiffusion with
    use_timesteps = all steps re-derives identical betas (respace.py:69-83)."""
    betas = cosine_betas(num_steps)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    # respace.py:76-81: new_betas = 1 - ac/last_ac  (numerically the same sequence, recomputed)
    last = 1.0
    nb = []
    for a in ac:
        nb.append(1 - a / last)
        last = a
    betas = np.array(nb, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return dict(
        betas=betas,
        alphas_cumprod=ac,
        sqrt_alphas_cumprod=np.sqrt(ac),
        sqrt_one_minus_alphas_cumprod=np.sqrt(1.0 - ac),
        posterior_variance=post_var,
        posterior_log_variance_clipped=np.log(np.append(post_var[1], post_var[1:])),
        posterior_mean_coef1=betas * np.sqrt(ac_prev) / (1.0 - ac),
        posterior_mean_coef2=(1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    )


def p_sample_update(tab: dict, x_t: torch.Tensor, x0: torch.Tensor, t: int, noise: torch.Tensor) -> torch.Tensor:
    """gaussian_diffusion.py:217-220 (posterior mean), :273-292 (FIXED_SMALL), :448-459 (p_sample);
    tables are cast to fp32 at lookup (`_extract_into_tensor`, :1275)."""
    c1 = torch.tensor(tab["posterior_mean_coef1"][t]).float()
    c2 = torch.tensor(tab["posterior_mean_coef2"][t]).float()
    logvar = torch.tensor(tab["posterior_log_variance_clipped"][t]).float()
    mean = c1 * x0 + c2 * x_t
    nonzero = 0.0 if t == 0 else 1.0
    return mean + nonzero * torch.exp(0.5 * logvar) * noise


# --------------------------------------------------------------------------------------------
# MF-MDM G denoiser forward (restated with the hoisted decomposition of SURVEY.md 8a')
# --------------------------------------------------------------------------------------------


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def encoder_layer(sd, p, x, nhead):
    """torch.nn.TransformerEncoderLayer, post-norm, batch_first=False, exact-erf GELU, eps 1e-5, no mask
    (constructed interaction_segment_mdm.py:63-70).  x: [S,B,d]."""
    S, B, d = x.shape
    hd = d // nhead
    qkv = F.linear(x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
    q, k, v = qkv.split(d, dim=-1)
    # [S,B,d] -> [B,H,S,hd]
    sh = lambda z: z.reshape(S, B, nhead, hd).permute(1, 2, 0, 3)
    q, k, v = sh(q), sh(k), sh(v)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    a = (att @ v).permute(2, 0, 1, 3).reshape(S, B, d)
    a = _lin(sd, p + "self_attn.out_proj", a)
    x = F.layer_norm(x + a, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    h = _lin(sd, p + "linear2", F.gelu(_lin(sd, p + "linear1", x)))
    x = F.layer_norm(x + h, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    return x


def hand_side_tokens(sd, hand_side):
    """interaction_segment_mdm.py:266-288: rh -> zeros, lh -> e0; anything else raises ValueError."""
    res = []
    for hs in hand_side:
        if hs == "rh":
            res.append(sd["hand_side_process.rh_embed"])
        elif hs == "lh":
            res.append(sd["hand_side_process.lh_embed"])
        else:
            raise ValueError(f"unexpected hand_side: {hs}")
    return torch.stack(res, 0)


def g_forward(sd: dict, cfg: dict, x: torch.Tensor, timesteps: torch.Tensor, batch: dict,
              text_feat: torch.Tensor) -> torch.Tensor:
    """`InterationSegmentMDM.forward` (interaction_segment_mdm.py:134-174) given the CLIP text feature
    `text_feat = encode_text(batch["text"])` ([B,clip_dim] fp32; :111-132 is library code kept in PyTorch).
    x: [B,99,1,T] -> [B,99,1,T]."""
    d, H = cfg["latent_dim"], cfg["num_heads"]
    B, nfeat, _, T = x.shape
    pe = sd["sequence_pos_encoder.pe"]  # [5000,1,d]
    # :142 embed_timestep  (TimestepEmbedder :214-215)
    emb_t = _lin(sd, "embed_timestep.time_embed.2", F.silu(_lin(sd, "embed_timestep.time_embed.0", pe[timesteps])))
    emb_t = emb_t.permute(1, 0, 2)  # [1,B,d]
    emb_text = _lin(sd, "embed_text", text_feat).reshape(1, B, d)  # :146 (mask_cond is identity in eval)
    emb_hs = hand_side_tokens(sd, batch["hand_side"]).unsqueeze(0)  # :150
    emb_shape = _lin(sd, "hand_shape_process.shape_embed", batch["shape"].mean(1)).unsqueeze(0)  # :300-301
    emb_obj = _lin(sd, "obj_embed_process.embedding", batch["obj_embedding"].mean(1)).unsqueeze(0)  # :260-261
    emb = torch.nan_to_num(torch.cat([emb_t, emb_text, emb_hs, emb_shape, emb_obj], 0))  # :157-158
    # :161 InputProcess (:224-229)
    hand = _lin(sd, "input_process.poseEmbedding", x.permute(3, 0, 1, 2).reshape(T, B, nfeat))
    # :162 ObjectInputProcess (:241-247): Linear on every (padded) object then mean over the padded axis
    obj = _lin(sd, "obj_input_process.poseEmbedding", batch["obj_traj"].permute(0, 2, 1, 3)).mean(2).permute(1, 0, 2)
    m = torch.cat((hand, obj), -1)
    xs = torch.nan_to_num(_lin(sd, "input_merge.2", F.silu(_lin(sd, "input_merge.0", m))))  # :164-166
    xseq = torch.cat((emb, xs), 0)
    xseq = xseq + pe[: xseq.shape[0]]  # :170 (dropout identity in eval)
    for l in range(cfg["num_layers"]):
        xseq = encoder_layer(sd, f"seqTransEncoder.layers.{l}.", xseq, H)
    out = xseq[emb.shape[0]:]
    out = _lin(sd, "output_process.poseFinal", out)  # :313-318
    out = out.reshape(T, B, nfeat, 1).permute(1, 2, 3, 0)
    return torch.nan_to_num(out)


def p_sample_loop(sd, cfg, batch, text_feat, shape, noise_fn, x_T=None, t_start=None, t_end=0, tab=None):
    """gaussian_diffusion.py:573-640 with clip_denoised=False, no cond_fn, skip_timesteps=0.
    `noise_fn(t, shape)` supplies eps for step t (the harness patches th.randn_like the same way)."""
    tab = tab or diffusion_tables(1000)
    B = shape[0]
    x = x_T.clone() if x_T is not None else noise_fn(1000, shape)
    t_start = 999 if t_start is None else t_start
    for t in range(t_start, t_end - 1, -1):
        ts = torch.full((B,), t, dtype=torch.long)
        x0 = g_forward(sd, cfg, x, ts, batch, text_feat)
        x = p_sample_update(tab, x, x0, t, noise_fn(t, shape))
    return x


# --------------------------------------------------------------------------------------------
# Rotation helpers
# --------------------------------------------------------------------------------------------


def rot6d_to_rotmat(d6: torch.Tensor) -> torch.Tensor:
    """src/dev_fn/transform/rotation.py:446-467 (b1,b2,b3 stacked as ROWS; F.normalize eps 1e-12)."""
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = F.normalize(b2, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-2)


def rotmat_to_quat(m: torch.Tensor) -> torch.Tensor:
    """rotation.py:167-213 (+ standardize_quat: w >= 0).  Best-conditioned candidate by argmax(q_abs)."""
    bd = m.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m.reshape(bd + (9,)), dim=-1)
    x = torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], -1)
    q_abs = torch.where(x > 0, torch.sqrt(torch.clamp(x, min=0)), torch.zeros_like(x))
    cand = torch.stack(
        [
            torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
            torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], -1),
            torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], -1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], -1),
        ],
        -2,
    )
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    sel = q_abs.argmax(-1)
    out = torch.gather(cand, -2, sel[..., None, None].expand(bd + (1, 4))).squeeze(-2)
    return torch.where(out[..., 0:1] < 0, -out, out)


def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    """thirdparty/manotorch/manotorch/utils/geometry.py:225-253 (un-normalised q, two_s = 2/|q|^2)."""
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack(
        (
            1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
            two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
            two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
        ),
        -1,
    )
    return o.reshape(q.shape[:-1] + (3, 3))


# --------------------------------------------------------------------------------------------
# ManoLayer FK (quat mode, center_idx=0)
# --------------------------------------------------------------------------------------------

_LEV1, _LEV2, _LEV3 = [1, 4, 7, 10, 13], [2, 5, 8, 11, 14], [3, 6, 9, 12, 15]
_JOINT_REORDER = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]


def mano_fk_from_rotmats(assets: dict, R: torch.Tensor, betas: torch.Tensor, side: str):
    """`ManoLayer.skinning_layer` (thirdparty/manotorch/manotorch/manolayer.py:128-266) with center_idx=0.
    R: [N,16,3,3], betas: [N,10] -> verts [N,778,3], joints [N,21,3] (root-centred)."""
    t = lambda k: torch.from_numpy(np.asarray(assets[k], dtype=np.float32))
    shapedirs, posedirs, v_t, Jreg, W = t("shapedirs"), t("posedirs"), t("v_template"), t("J_regressor"), t("weights")
    N = R.shape[0]
    B_S = torch.matmul(shapedirs, betas.transpose(1, 0)).permute(2, 0, 1)  # :139
    J = torch.matmul(Jreg, v_t[None] + B_S)  # :142  [N,16,3]
    rot_minus = (R[:, 1:] - torch.eye(3)).reshape(N, 135)  # :147-151
    B_P = torch.matmul(posedirs, rot_minus.transpose(0, 1)).permute(2, 0, 1)  # :154
    T_P = v_t[None] + B_S + B_P  # :157

    def hom(Rm, tv):  # [..,3,3],[..,3] -> [..,4,4]
        top = torch.cat([Rm, tv[..., None]], -1)
        bot = torch.tensor([0.0, 0.0, 0.0, 1.0]).expand(top.shape[:-2] + (1, 4))
        return torch.cat([top, bot], -2)

    G = [None] * 16
    G[0] = hom(R[:, 0], J[:, 0])  # :160-162
    for lev, par in ((_LEV1, None), (_LEV2, _LEV1), (_LEV3, _LEV2)):  # :164-193
        for n, k in enumerate(lev):
            p = 0 if par is None else par[n]
            G[k] = torch.matmul(G[p], hom(R[:, k], J[:, k] - J[:, p]))
    G = torch.stack(G, 1)  # [N,16,4,4] already in joint order (:195-198 reorder undone)
    Jh = torch.cat([J, J.new_zeros(N, 16, 1)], 2)
    tmp = torch.matmul(G, Jh.unsqueeze(3))  # :202-203
    Gp = G - torch.cat([tmp.new_zeros(N, 16, 4, 3), tmp], 3)  # :204
    Tm = torch.einsum("nkij,vk->nvij", Gp, W)  # :208
    Tp_h = torch.cat([T_P, torch.ones(N, T_P.shape[1], 1)], -1)
    verts = torch.einsum("nvij,nvj->nvi", Tm, Tp_h)[..., :3]  # :216-219
    joints = G[:, :, :3, 3]
    tips = verts[:, [745, 317, 444 if side == "right" else 445, 556, 673]]  # :224-227
    joints = torch.cat([joints, tips], 1)[:, _JOINT_REORDER]  # :229,240
    center = joints[:, 0:1]  # :242-249
    return verts - center, joints - center


def mano_fk_pose_repr(assets: dict, pose_repr: torch.Tensor, betas: torch.Tensor, side: str):
    """`batch_recover_mano_from_pose_repr` for one hand side (segment_refine_model.py:117-131):
    pose_repr [N,99] -> rot6d -> rotmat -> quat -> ManoLayer(quat) -> + tsl."""
    N = pose_repr.shape[0]
    tsl = pose_repr[:, 0:3]
    R = rot6d_to_rotmat(pose_repr[:, 3:99].reshape(N, 16, 6))
    q = rotmat_to_quat(R)
    v, j = mano_fk_from_rotmats(assets, quat_to_rotmat(q), betas, side)
    return v + tsl[:, None], j + tsl[:, None]


# --------------------------------------------------------------------------------------------
# Chamfer nearest neighbour (pytorch3d.ops.knn_points K=1 -- PARITY UNPINNED, semantics defined here)
# --------------------------------------------------------------------------------------------


def nn_query_numpy(x: np.ndarray, y: np.ndarray):
    """x [N,P1,3], y [N,P2,3] fp32 -> (d2 [N,P1] fp32, idx [N,P1] int64).
    Restates thirdparty/chamfer_distance/chamfer_distance/chamfer_distance.py:147,151,162 (K=1 squared L2,
    idx[...,-1]).  d = (dx*dx + dy*dy) + dz*dz, every op rounded to fp32 (numpy float32 never fuses),
    ties -> lowest index (np.argmin first-occurrence rule)."""
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    N, P1, _ = x.shape
    d2 = np.empty((N, P1), np.float32)
    idx = np.empty((N, P1), np.int64)
    for n in range(N):
        for s in range(0, P1, 128):
            xs = x[n, s:s + 128]
            dx = xs[:, None, 0] - y[n, None, :, 0]
            dy = xs[:, None, 1] - y[n, None, :, 1]
            dz = xs[:, None, 2] - y[n, None, :, 2]
            d = (dx * dx + dy * dy) + dz * dz
            i = np.argmin(d, axis=1)
            idx[n, s:s + 128] = i
            d2[n, s:s + 128] = d[np.arange(d.shape[0]), i]
    return d2, idx


_HERE = os.path.dirname(os.path.abspath(__file__))
_clib = None


def build_c_oracle(force: bool = False) -> str:
    """Compile oracle/nn_oracle.c with gcc (-ffp-contract=off so no FMA is ever formed)."""
    so = os.path.join(_HERE, "libtamf_oracle.so")
    src = os.path.join(_HERE, "nn_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                               "-fopenmp", "-o", so, src])
    return so


def _lib():
    global _clib
    if _clib is None:
        _clib = ctypes.CDLL(build_c_oracle())
        _clib.oracle_nn_query.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _clib.oracle_nn_query.restype = None
    return _clib


def nn_query(x: np.ndarray, y: np.ndarray, threads: int = 0):
    """Same contract as `nn_query_numpy`, plain C (oracle/nn_oracle.c); tests check the two agree bit for bit."""
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    N, P1, _ = x.shape
    P2 = y.shape[1]
    d2 = np.empty((N, P1), np.float32)
    idx = np.empty((N, P1), np.int64)
    _lib().oracle_nn_query(x.ctypes.data, y.ctypes.data, N, P1, P2, d2.ctypes.data, idx.ctypes.data, threads)
    return d2, idx


def vertex_normals(verts: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """pytorch3d `Meshes.verts_normals_packed` as recalled (UNPINNED): area-weighted face normals accumulated
    per vertex -- cross(v2-v1, v0-v1) added to v1, cross(v0-v2, v1-v2) to v2, cross(v1-v0, v2-v0) to v0 --
    then normalize(eps=1e-6).  verts [N,V,3], faces [F,3]."""
    v = np.asarray(verts, np.float32)
    f = np.asarray(faces, np.int64)
    out = np.zeros_like(v)
    v0, v1, v2 = v[:, f[:, 0]], v[:, f[:, 1]], v[:, f[:, 2]]
    for n in range(v.shape[0]):
        np.add.at(out[n], f[:, 1], np.cross(v2[n] - v1[n], v0[n] - v1[n]))
        np.add.at(out[n], f[:, 2], np.cross(v0[n] - v2[n], v1[n] - v2[n]))
        np.add.at(out[n], f[:, 0], np.cross(v1[n] - v0[n], v2[n] - v0[n]))
    nrm = np.sqrt((out * out).sum(-1, keepdims=True))
    return (out / np.maximum(nrm, 1e-6)).astype(np.float32)


def point2point_h2o(x: torch.Tensor, y: torch.Tensor):
    """The part of `point2point_signed` (model/loss/chamfer_distance.py:36-62) that R consumes:
    unsigned |x - y[idx_x]| (y_normals is None at every call site) and idx_x."""
    d2, idx = nn_query(x.numpy(), y.numpy())
    idx_t = torch.from_numpy(idx)
    near = y.gather(1, idx_t[..., None].expand(-1, -1, 3))
    return (x - near).norm(dim=2), idx_t


def point2point_signed(x: torch.Tensor, y: torch.Tensor, x_normals=None, y_normals=None):
    """`point2point_signed` in full (model/loss/chamfer_distance.py:4-64) over `ChamferDistance.forward`
    (chamfer_distance.py:147-162: knn_points both ways) -> (y2x_signed [N,P2], x2y_signed [N,P1], yidx_near [N,P2])."""
    N, P1, D = x.shape
    P2 = y.shape[1]
    _, ix = nn_query(x.numpy(), y.numpy())
    _, iy = nn_query(y.numpy(), x.numpy())
    ix, iy = torch.from_numpy(ix), torch.from_numpy(iy)
    xe, ye = ix.view(N, P1, 1).expand(N, P1, D), iy.view(N, P2, 1).expand(N, P2, D)
    x2y, y2x = x - y.gather(1, xe), y - x.gather(1, ye)
    y2x_signed, x2y_signed = y2x.norm(dim=2), x2y.norm(dim=2)
    if x_normals is not None:
        y_nn = x_normals.gather(1, ye)
        y2x_signed = y2x_signed * torch.bmm(y_nn.reshape(-1, 1, 3), y2x.reshape(-1, 3, 1)).reshape(N, -1).sign()
    if y_normals is not None:
        x_nn = y_normals.gather(1, xe)
        x2y_signed = x2y_signed * torch.bmm(x_nn.reshape(-1, 1, 3), x2y.reshape(-1, 3, 1)).reshape(N, -1).sign()
    return y2x_signed, x2y_signed, iy


def obj_world_points(obj_traj: torch.Tensor, pts: torch.Tensor) -> torch.Tensor:
    """`tslrot6d_to_transf` + `transf_point_array` (src/dev_fn/transform/transform.py:148-154, 36-53):
    obj_traj [T,9], pts [P,3] -> [T,P,3] = (R_t @ p^T)^T + t_t."""
    R = rot6d_to_rotmat(obj_traj[:, 3:9])
    return torch.matmul(R, pts.t()[None]).transpose(1, 2) + obj_traj[:, None, 0:3]


def h2o_dist(verts: torch.Tensor, obj_traj: torch.Tensor, obj_num, obj_points) -> torch.Tensor:
    """`multi_object_h2o_dist` (segment_refine_model.py:142-168): verts [B,T,778,3], obj_traj [B,nobj_max,T,9],
    obj_points list of np [nobj_b,P,3] -> [B,T,778]."""
    out = []
    for b in range(verts.shape[0]):
        pts = torch.from_numpy(np.asarray(obj_points[b], np.float32))
        world = torch.cat([obj_world_points(obj_traj[b, o], pts[o]) for o in range(int(obj_num[b]))], 1)
        out.append(point2point_h2o(verts[b], world)[0])
    return torch.stack(out, 0)


def recover_mano(assets_rh, assets_lh, pose_repr, shape, hand_side):
    """`batch_recover_mano_from_pose_repr` (segment_refine_model.py:107-140) without normals."""
    V, J = [], []
    for b in range(pose_repr.shape[0]):
        if hand_side[b] == "rh":
            v, j = mano_fk_pose_repr(assets_rh, pose_repr[b], shape[b], "right")
        elif hand_side[b] == "lh":
            v, j = mano_fk_pose_repr(assets_lh, pose_repr[b], shape[b], "left")
        else:
            raise ValueError(f"unexpected hand_side: {hand_side[b]}")
        V.append(v)
        J.append(j)
    return torch.stack(V, 0), torch.stack(J, 0)


def r_forward(sd: dict, cfg: dict, batch: dict, assets_rh: dict, assets_lh: dict, use_pc: bool = True,
              with_aux: bool = True) -> dict:
    """`SegmentRefineModel.forward` (segment_refine_model.py:170-250), normals omitted (they only feed the
    discarded y2x_signed, :165)."""
    d, H = cfg["latent_dim"], cfg["num_heads"]
    x_in = batch["sample_pose_repr"]
    B, T, _ = x_in.shape
    pe = sd["sequence_pos_encoder.pe"]
    emb = torch.nan_to_num(torch.cat([
        hand_side_tokens(sd, batch["hand_side"]).unsqueeze(0),
        _lin(sd, "hand_shape_process.shape_embed", batch["shape"].mean(1)).unsqueeze(0),
        _lin(sd, "obj_embed_process.embedding", batch["obj_embedding"].mean(1)).unsqueeze(0)], 0))
    hand = _lin(sd, "input_process.poseEmbedding", x_in.permute(1, 0, 2))
    obj = _lin(sd, "obj_input_process.poseEmbedding", batch["obj_traj"].permute(0, 2, 1, 3)).mean(2).permute(1, 0, 2)
    pts = batch["obj_pointcloud"] if use_pc else batch["obj_verts"]
    obj_num = [len(o) for o in batch["obj_list"]]
    sv, sj = recover_mano(assets_rh, assets_lh, x_in, batch["shape"], batch["hand_side"])
    s_h2o = h2o_dist(sv, batch["obj_traj"], obj_num, pts)
    dist_in = _lin(sd, "h2o_dist_input_process.poseEmbedding", s_h2o.permute(1, 0, 2))
    m = torch.cat((hand, obj, dist_in), -1)
    xs = torch.nan_to_num(_lin(sd, "input_merge.2", F.silu(_lin(sd, "input_merge.0", m))))
    xseq = torch.cat((emb, xs), 0)
    xseq = xseq + pe[: xseq.shape[0]]
    for l in range(cfg["num_layers"]):
        xseq = encoder_layer(sd, f"seqTransEncoder.layers.{l}.", xseq, H)
    out = _lin(sd, "output_process.poseFinal", xseq[3:]).permute(1, 0, 2)
    out = torch.nan_to_num(x_in + out)
    res = {"refine_pose_repr": out, "sample_hand_verts": sv, "sample_hand_joints": sj, "sample_h2o_dist": s_h2o}
    if with_aux:
        rv, rj = recover_mano(assets_rh, assets_lh, out, batch["shape"], batch["hand_side"])
        res.update(refine_hand_verts=rv, refine_hand_joints=rj,
                   refine_h2o_dist=h2o_dist(rv, batch["obj_traj"], obj_num, pts))
        tv, tj = recover_mano(assets_rh, assets_lh, batch["pose_repr"], batch["shape"], batch["hand_side"])
        res.update(target_hand_verts=tv, target_hand_joints=tj,
                   target_h2o_dist=h2o_dist(tv, batch["obj_traj"], obj_num, pts))
    return res
