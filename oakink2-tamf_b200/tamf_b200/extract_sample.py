"""G -> R sampling pipeline and the launchers' on-disk contract.

Reference call sites:
  * `extract_refined_sample(_bihand)`      src/oakink2_tamf/model/extract_sample.py:7-41, 44-110
  * `interaction_segment_collate`          src/oakink2_tamf/dataset/collate.py:6-59
  * `map_copy_select_to`                   src/dev_fn/transform/cast.py:76-85
  * sampling launcher loop + `.npy` layout src/oakink2_tamf/launch/sample.py:198-237
  * refine launcher loop + `save_dict.pkl` src/oakink2_tamf/launch/sample_refine.py:229-296
  * contact-ratio score helpers            script/compute_score/compute_score_cr.py:122-149

The reference runs ONE sequence per reverse chain (collate([gt_sample]) -> B = 1).  The batched forms here
(`extract_refined_samples`, `sample_dataset`, `refine_dataset`) stack up to `batch_size` dataset items into one chain
-- same per-item result layout, ~60x fewer kernel launches per item.  Items of one chain must share the frame count AND
the object count: the collate zero-pads `obj_traj` / `obj_embedding` to the largest object count of the batch and the
model averages over that padded axis (ObjectInputProcess / ObjectEmbedProcess, interaction_segment_mdm.py:243-246,
258-261), so an item batched with a larger one would see its object features scaled by nobj / nobj_max -- which the
B = 1 launchers never do.  Groups are therefore cut wherever T or obj_num changes, and each group draws its noise from
its own Philox stream (seed + first sample id).
"""
from __future__ import annotations

import os
import pickle
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch

from .chamfer import nn_query
from .shard import shard_range

# key classes of the reference collate (dataset/collate.py:6-30)
DEFAULT_COLLATE_KEY = ("pose_repr", "pose_repr_lh", "pose_repr_rh", "shape", "shape_lh", "shape_rh", "len", "mask",
                       "obj_num", "sample_pose_repr")
NO_COLLATE_KEY = ("hand_side", "text", "obj_list", "info", "obj_verts", "obj_faces", "obj_pointcloud", "sample_info",
                  "frame_id")
PAD_COLLATE_KEY = ("obj_traj", "obj_embedding")
SELECT_G = ("mask", "pose_repr", "shape", "obj_num", "obj_traj", "obj_embedding")
SELECT_R = SELECT_G + ("sample_pose_repr",)


def _stack(items):
    first = items[0]
    if isinstance(first, torch.Tensor):
        return torch.stack(items, 0)
    if isinstance(first, np.ndarray):
        return torch.from_numpy(np.stack(items, 0))
    if isinstance(first, (bool, np.bool_)):
        return torch.tensor(items, dtype=torch.bool)
    if isinstance(first, (int, np.integer)):
        return torch.tensor(items, dtype=torch.int64)
    if isinstance(first, (float, np.floating)):
        return torch.tensor(items, dtype=torch.float64)
    raise TypeError(f"cannot collate values of type {type(first)}")


def interaction_segment_collate(batch: Sequence[dict]) -> dict:
    """List of dataset items -> batch dict: stacked tensors, per-item lists, and the object axis of `obj_traj` /
    `obj_embedding` zero-padded to the largest object count of the batch (dataset/collate.py:33-59)."""
    res = {}
    for key in batch[0].keys():
        vals = [b[key] for b in batch]
        if key in DEFAULT_COLLATE_KEY:
            res[key] = _stack(vals)
        elif key in NO_COLLATE_KEY:
            res[key] = vals
        elif key in PAD_COLLATE_KEY:
            n = max(v.shape[0] for v in vals)
            padded = []
            for v in vals:
                v = np.asarray(v)
                if v.shape[0] < n:
                    v = np.concatenate((v, np.zeros((n - v.shape[0],) + v.shape[1:], v.dtype)), 0)
                padded.append(v)
            res[key] = _stack(padded)
        else:
            raise KeyError(f"unexpected key in batch! got {key}")
    return res


def map_copy_select_to(mapping: dict, device=None, dtype=None, select: Optional[Iterable[str]] = None) -> dict:
    """Shallow copy with the selected tensors moved / cast (dev_fn/transform/cast.py:76-85)."""
    sel = mapping if select is None else select
    return {k: (v.to(device=device, dtype=dtype) if k in sel else v) for k, v in mapping.items()}


def _g_chain(generation_model, diffusion, batch_device, seed=None):
    generation_model.eval()
    B, T, nf = batch_device["pose_repr"].shape
    with torch.no_grad():
        sample = diffusion.p_sample_loop(generation_model, (B, nf, 1, T), clip_denoised=False,
                                         model_kwargs={"batch": batch_device}, skip_timesteps=0, init_image=None,
                                         progress=False, dump_steps=None, noise=None, const_noise=False, seed=seed)
    return sample.permute((0, 3, 1, 2)).squeeze(3)  # [B,T,99]  (extract_sample.py:32)


def extract_refined_samples(generation_model, diffusion, refine_model, gt_samples: Sequence[dict], device,
                            dtype=torch.float32, seed=None) -> np.ndarray:
    """Batched `extract_refined_sample`: all items (same frame count) run as ONE G chain and ONE R pass.
    Returns refine_pose_repr [len(gt_samples), T, 99] as numpy."""
    if len({_item_obj_num(it) for it in gt_samples}) > 1:
        raise ValueError("extract_refined_samples: items of one batched chain must have the same object count (the "
                         "padded object axis is averaged over; use sample_dataset / one call per object count)")
    batch = interaction_segment_collate(list(gt_samples))
    batch_device = map_copy_select_to(batch, device=device, dtype=dtype, select=SELECT_G)
    batch_device["sample_pose_repr"] = _g_chain(generation_model, diffusion, batch_device, seed)
    refine_model.eval()
    with torch.no_grad():
        output = refine_model(batch_device)
    return output["refine_pose_repr"].detach().clone().cpu().numpy()


def extract_refined_sample(generation_model, diffusion, refine_model, gt_sample: dict, device, dtype=torch.float32,
                           seed=None) -> np.ndarray:
    """One dataset item -> refined pose representation [T,99] (extract_sample.py:7-41)."""
    return extract_refined_samples(generation_model, diffusion, refine_model, [gt_sample], device, dtype, seed)[0]


def bihand_item(gt_sample: dict, hand_side: str) -> dict:
    """The single-hand view of a two-hand dataset item: that side's pose / shape and only the objects paired with it
    (extract_sample.py:44-79; `obj_pair[1]` belongs to the right hand, `obj_pair[0]` to the left)."""
    if hand_side not in ("rh", "lh"):
        raise ValueError(f"unexpected hand_side: {hand_side}")
    pair = gt_sample["obj_pair"][1 if hand_side == "rh" else 0]
    obj_list = gt_sample["obj_list"]
    ids = [obj_list.index(o) for o in pair]
    return {
        "text": gt_sample["text"], "len": gt_sample["len"], "mask": gt_sample["mask"], "hand_side": hand_side,
        "pose_repr": gt_sample["pose_repr_rh" if hand_side == "rh" else "pose_repr_lh"],
        "shape": gt_sample["shape_rh" if hand_side == "rh" else "shape_lh"],
        "obj_num": len(pair), "obj_list": pair,
        "obj_traj": gt_sample["obj_traj"][ids, ...], "obj_embedding": gt_sample["obj_embedding"][ids, ...],
        "obj_pointcloud": gt_sample["obj_pointcloud"][ids, ...],
    }


def extract_refined_sample_bihand(generation_model, diffusion, refine_model, gt_sample: dict, hand_side: str, device,
                                  dtype=torch.float32, seed=None) -> np.ndarray:
    """extract_sample.py:44-110."""
    return extract_refined_sample(generation_model, diffusion, refine_model, bihand_item(gt_sample, hand_side), device,
                                  dtype, seed)


# ---- contact-ratio score helpers (script/compute_score/compute_score_cr.py:122-149) ----
def transf_merge_obj_pointcloud(obj_pointcloud: np.ndarray, obj_traj: np.ndarray) -> np.ndarray:
    """[nobj,nv,3] canonical clouds, [nobj,T,9] (tsl3 + rot6d) -> world points [T, nobj*nv, 3]."""
    pc = torch.as_tensor(obj_pointcloud, dtype=torch.float32)
    tr = torch.as_tensor(obj_traj, dtype=torch.float32)
    a1, a2 = tr[..., 3:6], tr[..., 6:9]
    b1 = torch.nn.functional.normalize(a1, dim=-1)
    b2 = torch.nn.functional.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    R = torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), dim=-2)  # rows b1, b2, b3 (rotation.py:446-467)
    world = torch.einsum("otij,onj->otni", R, pc) + tr[..., None, 0:3]  # [nobj,T,nv,3]
    return world.permute(1, 0, 2, 3).reshape(tr.shape[1], -1, 3).numpy()


def contact_min_cdist(hv, pc, device, dtype=torch.float32) -> List[float]:
    """Per-frame minimum hand-vertex to object-point distance: the reference builds the full torch.cdist matrix
    [T,778,P] and reduces it; here it is the minimum over the exact nearest-neighbour distances (tamf_nn_query)."""
    hv_t = torch.as_tensor(hv).to(device=device, dtype=torch.float32).contiguous()
    pc_t = torch.as_tensor(pc).to(device=device, dtype=torch.float32).contiguous()
    d2, _ = nn_query(hv_t, pc_t)
    return torch.sqrt(d2.min(dim=1).values).to(dtype).cpu().numpy().tolist()


# ---- launcher loops ----
def _item_obj_num(item: dict) -> int:
    return int(np.asarray(item["obj_traj"]).shape[0])


def _same_len_batches(dataset, ids: range, batch_size: int):
    """Consecutive ids, cut where the frame count or the object count changes or `batch_size` is reached (so that a
    batched chain is exactly the stack of the reference's B = 1 chains: no zero-padded object rows inside a group)."""
    cur, cur_key = [], None
    for i in ids:
        item = dataset[i]
        key = (int(np.asarray(item["pose_repr"]).shape[0]), _item_obj_num(item))
        if cur and (key != cur_key or len(cur) == batch_size):
            yield cur
            cur = []
        cur.append((i, item))
        cur_key = key
    if cur:
        yield cur


def _group_seed(seed, group):
    """A distinct Philox stream per group: with one seed for all, every group would replay the same per-step noise."""
    return None if seed is None else int(seed) + int(group[0][0])


def sample_dataset(model, diffusion, dataset, out_dir: Optional[str], worker_id: int = 0, num_worker: int = 1,
                   batch_size: int = 64, device=None, dtype=torch.float32, commit: bool = True, seed=None) -> dict:
    """The loop of `launch/sample.py:198-237`: worker `worker_id` of `num_worker` samples its contiguous share of the
    dataset and writes one `%06d.npy` ([T,99] fp32) per item under `out_dir` (= <ckpt>/sample/<save_offset>).
    Returns {sample_id: array}."""
    device = device if device is not None else next(model.parameters()).device
    out = {}
    for group in _same_len_batches(dataset, shard_range(len(dataset), worker_id, num_worker), batch_size):
        batch = interaction_segment_collate([it for _, it in group])
        batch_device = map_copy_select_to(batch, device=device, dtype=dtype, select=SELECT_G)
        arr = _g_chain(model, diffusion, batch_device, _group_seed(seed, group)).detach().cpu().numpy()
        for (sid, _), a in zip(group, arr):
            out[sid] = a
            if commit and out_dir is not None:
                os.makedirs(out_dir, exist_ok=True)
                np.save(os.path.join(out_dir, f"{sid:06d}.npy"), a)
    return out


def refine_dataset(refine_model, dataset, out_dir: Optional[str], worker_id: int = 0, num_worker: int = 1,
                   batch_size: int = 64, device=None, dtype=torch.float32, commit: bool = True) -> List[dict]:
    """The loop of `launch/sample_refine.py:229-296`: R forward on items that carry `sample_pose_repr`, then one
    `save_dict.pkl` per item under out_dir/<process_key with '/' -> '++'>/<info[1]>/<info[2]>/ with the reference's
    keys.  joints / verts are the FK of the refined pose (already part of the R output dict)."""
    device = device if device is not None else next(refine_model.parameters()).device
    refine_model.eval()
    saved, seen = [], set()
    for group in _same_len_batches(dataset, shard_range(len(dataset), worker_id, num_worker), batch_size):
        group = [(i, it) for i, it in group if not (it["info"] in seen or seen.add(it["info"]))]  # duplicate_check
        if not group:
            continue
        batch = interaction_segment_collate([it for _, it in group])
        batch_device = map_copy_select_to(batch, device=device, dtype=dtype, select=SELECT_R)
        with torch.no_grad():
            output = refine_model(batch_device)
        pose = output["refine_pose_repr"].detach().cpu().numpy()
        joints = output["refine_hand_joints"].detach().cpu().numpy()
        verts = output["refine_hand_verts"].detach().cpu().numpy()
        for k, (sid, it) in enumerate(group):
            side = it["hand_side"]
            layer = refine_model.mano_layer_rh if side == "rh" else refine_model.mano_layer_lh
            info = it["info"]
            d = {"process_key": info[0], "info": info, "hand_side": side, "joints": joints[k], "verts": verts[k],
                 "faces": layer.get_mano_closed_faces().cpu().numpy(), "obj_list": it["obj_list"], "len": it["len"],
                 "frame_id": it.get("frame_id"), "refine_pose_repr": pose[k]}
            saved.append(d)
            if commit and out_dir is not None:
                path = os.path.join(out_dir, str(info[0]).replace("/", "++"), str(info[1]), str(info[2]), "save_dict.pkl")
                os.makedirs(os.path.dirname(path), exist_ok=True)
                with open(path, "wb") as f:
                    pickle.dump(d, f)
    return saved
