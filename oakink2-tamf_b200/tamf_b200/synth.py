"""Seeded synthetic inputs, MANO-shaped assets and random-init weights (numpy only, so they travel).

The dataset, MANO .pkl and checkpoints of the reference are not available offline (SURVEY.md 8d
"Synthetic inputs"), so tests, bench and smoke all draw from these generators.  Shapes follow the
reference's batch schema (dataset/interaction_segment.py:415-448, dataset/collate.py:6-30).
"""
from __future__ import annotations

import numpy as np
import torch

ARCH = {
    # config/arch_mdm.yml, config/arch_mdm_l.yml, config/arch_refine.yml (read verbatim; activation gelu)
    "arch_mdm": dict(input_dim=99, obj_input_dim=9, hand_shape_dim=10, obj_embed_dim=768, latent_dim=256,
                     ff_size=1024, num_layers=8, num_heads=4, dropout=0.1, activation="gelu"),
    "arch_mdm_l": dict(input_dim=99, obj_input_dim=9, hand_shape_dim=10, obj_embed_dim=768, latent_dim=512,
                       ff_size=2048, num_layers=8, num_heads=4, dropout=0.1, activation="gelu"),
    "arch_refine": dict(input_dim=99, obj_input_dim=9, hand_shape_dim=10, obj_embed_dim=768, latent_dim=256,
                        ff_size=1024, num_layers=8, num_heads=4, dropout=0.1, activation="gelu"),
}

N_VERTS, N_FACES, N_JOINTS = 778, 1538, 16


def mano_assets(side: str = "right", seed: int = 7) -> dict:
    """MANO-shaped synthetic assets: same shapes/dtypes as manolayer.py:75-81 buffers."""
    rng = np.random.default_rng(seed + (0 if side == "right" else 1000))
    v_template = rng.normal(0, 0.03, (N_VERTS, 3))
    shapedirs = rng.normal(0, 0.002, (N_VERTS, 3, 10))
    posedirs = rng.normal(0, 0.001, (N_VERTS, 3, 135))
    w = rng.random((N_VERTS, N_JOINTS)) ** 8
    weights = w / w.sum(1, keepdims=True)
    J = np.zeros((N_JOINTS, N_VERTS))
    for j in range(N_JOINTS):
        nz = rng.choice(N_VERTS, 12, replace=False)
        v = rng.random(12)
        J[j, nz] = v / v.sum()
    faces = rng.integers(0, N_VERTS, (N_FACES, 3)).astype(np.int64)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(v_template=f32(v_template), shapedirs=f32(shapedirs), posedirs=f32(posedirs),
                weights=f32(weights), J_regressor=f32(J), faces=faces)


def _linear(rng, out_f, in_f, prefix, sd, scale=1.0):
    k = scale / np.sqrt(in_f)
    sd[prefix + ".weight"] = rng.uniform(-k, k, (out_f, in_f)).astype(np.float32)
    sd[prefix + ".bias"] = rng.uniform(-k, k, (out_f,)).astype(np.float32)


def positional_table(d: int, max_len: int = 5000) -> np.ndarray:
    """PositionalEncoding.pe (interaction_segment_mdm.py:186-193) computed with torch fp32 ops, [max_len,1,d]."""
    pe = torch.zeros(max_len, d)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d, 2).float() * (-np.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).transpose(0, 1).contiguous().numpy()


def _encoder_layers(rng, sd, d, ff, L):
    for l in range(L):
        p = f"seqTransEncoder.layers.{l}."
        k = 1.0 / np.sqrt(d)
        sd[p + "self_attn.in_proj_weight"] = rng.uniform(-k, k, (3 * d, d)).astype(np.float32)
        sd[p + "self_attn.in_proj_bias"] = rng.uniform(-0.05, 0.05, (3 * d,)).astype(np.float32)
        _linear(rng, d, d, p + "self_attn.out_proj", sd)
        _linear(rng, ff, d, p + "linear1", sd)
        _linear(rng, d, ff, p + "linear2", sd)
        for n in ("norm1", "norm2"):
            sd[p + n + ".weight"] = (1.0 + 0.1 * rng.standard_normal(d)).astype(np.float32)
            sd[p + n + ".bias"] = (0.05 * rng.standard_normal(d)).astype(np.float32)


def g_state_dict(cfg: dict, seed: int = 0) -> dict:
    """Random-init weights for InterationSegmentMDM with the reference's key names,
    clip_model.* excluded exactly like saved checkpoints (util/state_util.py:22-39)."""
    rng = np.random.default_rng(seed)
    d, ff, L = cfg["latent_dim"], cfg["ff_size"], cfg["num_layers"]
    sd: dict = {}
    _linear(rng, d, cfg["hand_shape_dim"], "hand_shape_process.shape_embed", sd)
    _linear(rng, d, cfg["obj_embed_dim"], "obj_embed_process.embedding", sd)
    _linear(rng, d, cfg["input_dim"], "input_process.poseEmbedding", sd)
    _linear(rng, d, cfg["obj_input_dim"], "obj_input_process.poseEmbedding", sd)
    _linear(rng, d, 2 * d, "input_merge.0", sd)
    _linear(rng, d, d, "input_merge.2", sd)
    _encoder_layers(rng, sd, d, ff, L)
    _linear(rng, d, d, "embed_timestep.time_embed.0", sd)
    _linear(rng, d, d, "embed_timestep.time_embed.2", sd)
    _linear(rng, d, cfg.get("clip_dim", 512), "embed_text", sd)
    _linear(rng, cfg["input_dim"], d, "output_process.poseFinal", sd)
    rh = np.zeros(d, np.float32)
    lh = np.zeros(d, np.float32)
    lh[0] = 1.0
    sd["hand_side_process.rh_embed"] = rh
    sd["hand_side_process.lh_embed"] = lh
    sd["sequence_pos_encoder.pe"] = positional_table(d)
    sd["embed_timestep.sequence_pos_encoder.pe"] = sd["sequence_pos_encoder.pe"]  # shared module, aliased key
    return {k: torch.from_numpy(v) for k, v in sd.items()}


def r_state_dict(cfg: dict, seed: int = 0) -> dict:
    """Random-init weights for SegmentRefineModel (segment_refine_model.py:21-105); mano buffers excluded."""
    rng = np.random.default_rng(seed + 77)
    d, ff, L = cfg["latent_dim"], cfg["ff_size"], cfg["num_layers"]
    sd: dict = {}
    _linear(rng, d, cfg["hand_shape_dim"], "hand_shape_process.shape_embed", sd)
    _linear(rng, d, cfg["obj_embed_dim"], "obj_embed_process.embedding", sd)
    _linear(rng, d, cfg["input_dim"], "input_process.poseEmbedding", sd)
    _linear(rng, d, cfg["obj_input_dim"], "obj_input_process.poseEmbedding", sd)
    _linear(rng, d, 778, "h2o_dist_input_process.poseEmbedding", sd)
    _linear(rng, d, 3 * d, "input_merge.0", sd)
    _linear(rng, d, d, "input_merge.2", sd)
    _encoder_layers(rng, sd, d, ff, L)
    # small output head => residual refinement stays near the input pose (keeps FK well conditioned)
    _linear(rng, cfg["input_dim"], d, "output_process.poseFinal", sd, scale=0.05)
    rh = np.zeros(d, np.float32)
    lh = np.zeros(d, np.float32)
    lh[0] = 1.0
    sd["hand_side_process.rh_embed"] = rh
    sd["hand_side_process.lh_embed"] = lh
    sd["sequence_pos_encoder.pe"] = positional_table(d)
    return {k: torch.from_numpy(v) for k, v in sd.items()}


TEXTS = [
    "pick up the bottle and pour water into the cup",
    "open the box with both hands",
    "stir the bowl with a spoon",
    "cut the apple with the knife",
]


def random_pose_repr(rng, B, T):
    """[B,T,99]: tsl ~ 0.1 N(0,1); 16 x rot6d near identity with noise (valid input to Gram-Schmidt)."""
    tsl = 0.1 * rng.standard_normal((B, T, 3))
    r6 = np.tile(np.array([1, 0, 0, 0, 1, 0], np.float64), (B, T, 16, 1)) + 0.4 * rng.standard_normal((B, T, 16, 6))
    return np.concatenate([tsl, r6.reshape(B, T, 96)], -1).astype(np.float32)


def make_batch(B: int, T: int = 160, nobj: int = 2, seed: int = 0, ragged: bool = False,
               npoints: int = 8192, with_pointcloud: bool = False) -> dict:
    """A collated batch dict in the reference's schema.  `ragged` zero-pads the object axis
    (nobj_b in {1..nobj}) exactly like `interaction_segment_collate` (dataset/collate.py:41-53)."""
    rng = np.random.default_rng(seed)
    shape = np.repeat(0.5 * rng.standard_normal((B, 1, 10)), T, axis=1).astype(np.float32)
    obj_traj = np.concatenate(
        [0.1 * rng.standard_normal((B, nobj, T, 3)), rng.standard_normal((B, nobj, T, 6))], -1).astype(np.float32)
    obj_emb = rng.standard_normal((B, nobj, 768)).astype(np.float32)
    obj_num = np.full((B,), nobj, np.int64)
    if ragged:
        for b in range(B):
            n = 1 + (b % nobj)
            obj_num[b] = n
            obj_traj[b, n:] = 0.0
            obj_emb[b, n:] = 0.0
    batch = {
        "text": [TEXTS[b % len(TEXTS)] for b in range(B)],
        "hand_side": ["rh" if b % 2 == 0 else "lh" for b in range(B)],
        "shape": torch.from_numpy(shape),
        "obj_traj": torch.from_numpy(obj_traj),
        "obj_embedding": torch.from_numpy(obj_emb),
        "obj_num": torch.from_numpy(obj_num),
        "pose_repr": torch.from_numpy(random_pose_repr(rng, B, T)),
        "mask": torch.ones(B, T, dtype=torch.bool),
        "len": torch.full((B,), T, dtype=torch.int64),
        "obj_list": [[f"obj_{b}_{i}" for i in range(int(obj_num[b]))] for b in range(B)],
    }
    if with_pointcloud:
        batch["obj_pointcloud"] = [
            (0.05 * rng.standard_normal((int(obj_num[b]), npoints, 3))).astype(np.float32) for b in range(B)]
        sig = rng.uniform(0.02, 0.1, (B, 1, 1))
        pr = batch["pose_repr"].numpy().copy()
        pr[..., :3] += (0.1 * sig * rng.standard_normal((B, T, 3))).astype(np.float32)
        pr[..., 3:] += (sig * rng.standard_normal((B, T, 96))).astype(np.float32)
        batch["sample_pose_repr"] = torch.from_numpy(pr.astype(np.float32))
    return batch


def make_items(n: int, T: int = 160, nobj: int = 2, seed: int = 0, ragged: bool = True, npoints: int = 256,
               bihand: bool = False) -> list:
    """`n` dataset items in the schema of `InteractionSegmentData.__getitem__` (dataset/interaction_segment.py:415-448):
    numpy arrays WITHOUT the batch axis and with the item's own object count (the collate pads).  `bihand` adds the
    two-hand keys `extract_refined_sample_bihand` reads (pose_repr_{lh,rh}, shape_{lh,rh}, obj_pair)."""
    b = make_batch(n, T, nobj=nobj, seed=seed, ragged=ragged, npoints=npoints, with_pointcloud=True)
    rng = np.random.default_rng(seed + 12345)
    items = []
    for i in range(n):
        k = int(b["obj_num"][i])
        it = {
            "text": b["text"][i], "hand_side": b["hand_side"][i], "shape": b["shape"][i].numpy(),
            "obj_traj": b["obj_traj"][i, :k].numpy(), "obj_embedding": b["obj_embedding"][i, :k].numpy(), "obj_num": k,
            "pose_repr": b["pose_repr"][i].numpy(), "sample_pose_repr": b["sample_pose_repr"][i].numpy(),
            "mask": b["mask"][i].numpy(), "len": T, "obj_list": list(b["obj_list"][i]),
            "obj_pointcloud": b["obj_pointcloud"][i], "info": (f"scene/{i:03d}", i, 0),
            "frame_id": list(range(T)),
        }
        if bihand:
            it["pose_repr_rh"], it["shape_rh"] = it["pose_repr"], it["shape"]
            it["pose_repr_lh"] = random_pose_repr(rng, 1, T)[0]
            it["shape_lh"] = np.repeat(0.5 * rng.standard_normal((1, 10)), T, axis=0).astype(np.float32)
            it["obj_pair"] = (it["obj_list"][:1], it["obj_list"][-1:])  # [0]: left-hand objects, [1]: right-hand
        items.append(it)
    return items


def text_features(texts, dim: int = 512, seed: int = 99) -> torch.Tensor:
    """Deterministic stand-in for CLIP `encode_text` output ([B,512] fp32): a hash-seeded unit-scale vector per
    string.  CLIP weights are a network download (not offline), so benches/tests use this synthetic encoder."""
    out = np.empty((len(texts), dim), np.float32)
    for i, s in enumerate(texts):
        h = (hash_str(s) + seed) % (2 ** 32)
        out[i] = np.random.default_rng(h).standard_normal(dim).astype(np.float32) * 0.3
    return torch.from_numpy(out)


def hash_str(s: str) -> int:
    h = 2166136261
    for c in s.encode("utf-8"):
        h = ((h ^ c) * 16777619) % (2 ** 32)
    return h


def step_noise(seed: int, t: int, shape) -> torch.Tensor:
    """Counter-based per-step noise eps_t = randn(seed (+) t) shared by oracle and device path."""
    g = torch.Generator().manual_seed((seed * 1000003 + t) % (2 ** 63))
    return torch.randn(*shape, generator=g)


def p2p_clouds(seed: int = 13, T: int = 2, nobj: int = 2, P: int = 8192):
    """Inputs of `point2point_signed` at the reference's call shape (segment_refine_model.py:165,
    interaction_segment_extra_loss.py:157): x = hand-sized cloud [T,778,3] with unit normals, y = nobj * P object points
    [T, nobj*P, 3] with unit normals (numpy fp32)."""
    rng = np.random.default_rng(seed)
    x = (0.08 * rng.standard_normal((T, 778, 3))).astype(np.float32)
    xn = rng.standard_normal((T, 778, 3)).astype(np.float32)
    xn /= np.linalg.norm(xn, axis=-1, keepdims=True)
    y = (0.1 * rng.standard_normal((T, nobj * P, 3))).astype(np.float32)
    yn = rng.standard_normal((T, nobj * P, 3)).astype(np.float32)
    yn /= np.linalg.norm(yn, axis=-1, keepdims=True)
    return x, xn, y, yn
