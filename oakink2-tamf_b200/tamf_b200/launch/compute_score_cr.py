"""python -m tamf_b200.launch.compute_score_cr -- contact ratio (CR) of refined samples against the ground truth.

The reference's script/compute_score/compute_score_cr.py (reg_entry :52-110, main :152-300): for every distinct item the
refined hand vertices are read from <sample_refine_filepath>/<process_key with '/' -> '++'>/<info[1]>/<info[2]>/
save_dict.pkl (what launch/sample_refine writes), the ground-truth vertices come from ManoLayer FK of the item's
pose_repr, both are cut to the item's available length, the objects' canonical clouds are moved by the object
trajectory and merged, and the per-frame minimum hand-vertex to object-point distance decides contact (< 5 mm).  Prints
the two array shapes and the two ratios, saves ./tmp/compute_score/contact_ratio/{gt,refined}_contact_dist.npy.

The reference builds the [T,778,P] torch.cdist matrix per item; here the minimum comes from the exact nearest-neighbour
kernel (tamf_nn_query) and FK from tamf_mano_fk.  The device the reference hard-codes (`cuda:4`, :185) is
--runtime.device_id."""
from __future__ import annotations

import logging
import os
import pickle
from typing import List

import numpy as np
import torch

from . import config as C
from .data import open_dataset

PROG = "compute_score_cr"
WS_DIR = os.getcwd()
CONTACT_THRESHOLD = 0.005  # metres (compute_score_cr.py:289-290)
_logger = logging.getLogger(__name__)


def reg_entry(reg: C.Registry) -> None:
    reg.register("data_prefix", prefix="data", category=str, default=f"{WS_DIR}/data", abspath=True, required=True)
    reg.register("process_range", prefix="data", category=List[str], seq=":",
                 default=[f"?(file:{WS_DIR}/asset/split/test.txt)"])
    reg.register("obj_embedding_prefix", prefix="data", category=str, abspath=True,
                 default="common/retrieve_obj_embedding/main/embedding")
    reg.register("obj_pointcloud_prefix", prefix="data", category=str, abspath=True,
                 default="common/retrieve_obj_pointcloud/main/pointcloud")
    reg.register("cache_dict_filepath", prefix="data", category=str, abspath=True,
                 default="common/save_cache_dict/main/cache/test.pkl")
    reg.register("source", prefix="data", category=str, default=None,
                 desc="reference | items:FILE.pkl | synthetic:N[:T[:K]] (launch/data.py)")
    C.reg_mano_param(reg, "mano", WS_DIR)
    C.reg_model_param(reg, "model")
    reg.register("sample_refine_filepath", prefix="debug", category=str, abspath=True,
                 default=f"{WS_DIR}/common/sample_refine/main/sample/test/arch_mdm_l__0399")
    reg.register("synthetic_mano", prefix="debug", category=bool, default=False,
                 desc="synthetic MANO-shaped assets instead of mano.mano_path (dry runs)")
    reg.register("out_dir", prefix="debug", category=str, default="./tmp/compute_score/contact_ratio/", abspath=True)
    reg.register("device_id", prefix="runtime", category=int, default=4)


def reg_extract(reg: C.Registry) -> dict:
    return {p: reg.select(p) for p in ("data", "debug", "mano", "model", "runtime")}


def contact_distances(dataset, sample_refine_filepath: str, layers: dict, device) -> tuple:
    """-> (gt_contact_dist, refined_contact_dist): per-frame minima over every distinct item that has a save_dict.pkl."""
    from .. import contact_min_cdist, transf_merge_obj_pointcloud
    gt_all, refined_all, seen = [], [], set()
    for i in range(len(dataset)):
        item = dataset[i]
        info = item["info"]
        if info in seen:  # duplicate_check (:212-217)
            continue
        seen.add(info)
        path = os.path.join(sample_refine_filepath, str(info[0]).replace("/", "++"), str(info[1]), str(info[2]),
                            "save_dict.pkl")
        if not os.path.exists(path):
            continue
        with open(path, "rb") as f:
            refined_verts = pickle.load(f)["verts"]
        side, n = item["hand_side"], int(item["len"])
        if side not in layers:
            raise ValueError(f"unexpected hand_side: {side}")
        pose = torch.as_tensor(np.asarray(item["pose_repr"]), dtype=torch.float32, device=device)
        betas = torch.as_tensor(np.asarray(item["shape"]), dtype=torch.float32, device=device)
        gt_verts, _ = layers[side].forward_pose_repr(pose, betas)  # FK of rot6d -> quat + translation (:246-262)
        gt_verts = gt_verts.cpu().numpy()[:n]
        merged = transf_merge_obj_pointcloud(np.asarray(item["obj_pointcloud"]), np.asarray(item["obj_traj"])[:, :n])
        gt_all.extend(contact_min_cdist(gt_verts, merged, device))
        refined_all.extend(contact_min_cdist(np.asarray(refined_verts)[:n], merged, device))
    return np.array(gt_all), np.array(refined_all)


def main(argv=None) -> tuple:
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s %(message)s")
    from .. import ManoLayer, synth
    reg = C.Registry(PROG)
    reg_entry(reg)
    reg.parse(argv)
    run_cfg = reg_extract(reg)
    _logger.info("run_cfg: %s", run_cfg)
    dataset = open_dataset(run_cfg["data"], enable_obj_model=True, with_pointcloud=True)
    device = torch.device(f"cuda:{run_cfg['runtime']['device_id']}")
    torch.cuda.set_device(device)
    mk = lambda side: ManoLayer(mano_assets_root=run_cfg["mano"]["mano_path"], rot_mode="quat", side=side, center_idx=0,
                                use_pca=False, flat_hand_mean=True,
                                assets=synth.mano_assets(side) if run_cfg["debug"]["synthetic_mano"] else None).to(device)
    layers = {"rh": mk("right"), "lh": mk("left")}
    gt, refined = contact_distances(dataset, run_cfg["debug"]["sample_refine_filepath"], layers, device)
    print(gt.shape, refined.shape)
    gt_ratio = float(np.mean(gt < CONTACT_THRESHOLD)) if gt.size else float("nan")
    refined_ratio = float(np.mean(refined < CONTACT_THRESHOLD)) if refined.size else float("nan")
    print(gt_ratio, refined_ratio)
    out = run_cfg["debug"]["out_dir"]
    os.makedirs(out, exist_ok=True)
    np.save(os.path.join(out, "gt_contact_dist.npy"), gt)
    np.save(os.path.join(out, "refined_contact_dist.npy"), refined)
    return gt_ratio, refined_ratio


if __name__ == "__main__":
    main()
