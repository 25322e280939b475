"""python -m tamf_b200.launch.sample_refine -- refine generated samples with MF-MDM R and save hand meshes.

Entry point behind the reference's script/sample_refine.sh:

    python -m tamf_b200.launch.sample_refine --data.process_range "?(file:./asset/split/test.txt)" \\
        --data.cache_dict_filepath common/.../test.pkl --debug.model_weight_filepath refine.pt \\
        --debug.sample_save_offset test/arch_refine__0399 --commit

Same flags and output layout as src/oakink2_tamf/launch/sample_refine.py (reg_entry :48-117, main :131-296): every
distinct item (duplicate `info` skipped, :231-236) goes through SegmentRefineModel and, with --commit, one
<ckpt_path>/sample/<sample_save_offset>/<process_key with '/' -> '++'>/<info[1]>/<info[2]>/save_dict.pkl holding
process_key, info, hand_side, joints, verts, faces (closed), obj_list, len, frame_id, refine_pose_repr is written.
The reference hard-codes two things this launcher exposes as flags with those values as defaults: the directory of
generated samples (`common/sample/main/sample/test/arch_mdm_l__0399`, :172) as --data.sample_dir and the device
(`cuda:4`, :174) as --runtime.device_id.  Items of equal frame and object count are refined in batches of
--runtime.batch_size (the reference: one at a time); the model config defaults to arch_refine (config/arch_refine.yml
== launch/param/model.py defaults)."""
from __future__ import annotations

import logging
import os
from typing import List

import torch

from . import config as C
from .data import GeneratedPoseReprSamples, open_dataset

PROG = "sample_refine"
WS_DIR = os.getcwd()
_logger = logging.getLogger(__name__)


def reg_entry(reg: C.Registry) -> None:
    reg.register("data_prefix", prefix="data", category=str, default=f"{WS_DIR}/data", abspath=True, required=True)
    reg.register("process_range", prefix="data", category=List[str], seq=":",
                 default=[f"?(file:{WS_DIR}/mocap_meta/process_range/test.txt)"])
    reg.register("obj_embedding_prefix", prefix="data", category=str, abspath=True,
                 default="common/retrieve_obj_embedding/main/embedding")
    reg.register("obj_pointcloud_prefix", prefix="data", category=str, abspath=True,
                 default="common/retrieve_obj_pointcloud/main/pointcloud")
    reg.register("cache_dict_filepath", prefix="data", category=str, abspath=True,
                 default="common/save_cache_dict/main/cache/test.pkl")
    reg.register("source", prefix="data", category=str, default=None,
                 desc="reference | items:FILE.pkl | synthetic:N[:T[:K]] (launch/data.py)")
    reg.register("sample_dir", prefix="data", category=List[str], seq=":",
                 default=["common/sample/main/sample/test/arch_mdm_l__0399"],
                 desc="directories of generated %06d.npy samples; '-' = the items already carry sample_pose_repr")
    C.reg_mano_param(reg, "mano", WS_DIR)
    C.reg_model_param(reg, "model")
    reg.register("model_weight_filepath", prefix="debug", category=str, abspath=True)
    reg.register("sample_save_offset", prefix="debug", category=str)
    reg.register("random_init_seed", prefix="debug", category=int, default=None,
                 desc="no checkpoint: random weights (and synthetic MANO assets) from this seed (dry runs)")
    reg.register("device_id", prefix="runtime", category=int, default=4)
    reg.register("batch_size", prefix="runtime", category=int, default=64)


def reg_extract(reg: C.Registry) -> dict:
    return {p: reg.select(p) for p in ("data", "debug", "mano", "model", "runtime")}


def build_model(run_cfg: dict, device: torch.device):
    from .. import SegmentRefineModel, synth
    mc = run_cfg["model"]
    kw = {k: mc[k] for k in C.MODEL_DEFAULTS}
    use_pc = run_cfg["data"].get("obj_pointcloud_prefix") is not None  # launch/sample_refine.py:213
    path, seed = run_cfg["debug"].get("model_weight_filepath"), run_cfg["debug"].get("random_init_seed")
    if path:
        model = SegmentRefineModel(run_cfg["mano"]["mano_path"], **kw, use_pc=use_pc)
        state_dict = torch.load(path, map_location="cpu")
    elif seed is not None:
        assets = {"right": synth.mano_assets("right"), "left": synth.mano_assets("left")}
        model = SegmentRefineModel(None, **kw, use_pc=use_pc, mano_assets=assets)
        state_dict = synth.r_state_dict(dict(mc), seed=int(seed))
    else:
        raise SystemExit("--debug.model_weight_filepath (or --debug.random_init_seed for a dry run) is required")
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    missing = [k for k in missing if not k.startswith("clip_model")]
    return model.to(device).eval(), missing, unexpected


def main(argv=None) -> list:
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s %(message)s")
    from .. import refine_dataset
    reg = C.Registry(PROG)
    C.reg_ckpt(reg, exp_id_default="main")
    reg_entry(reg)
    reg.parse(argv)
    ckpt_cfg, run_cfg = C.ckpt_extract(reg), reg_extract(reg)
    C.ckpt_setup(ckpt_cfg, _logger)
    C.ckpt_opt(ckpt_cfg, ckpt=ckpt_cfg, run=run_cfg)
    _logger.info("run_cfg: %s", run_cfg)
    dataset = open_dataset(run_cfg["data"], enable_obj_model=True, with_pointcloud=True)
    if run_cfg["data"]["sample_dir"] != ["-"]:
        dataset = GeneratedPoseReprSamples(dataset, run_cfg["data"]["sample_dir"])
    device = torch.device(f"cuda:{run_cfg['runtime']['device_id']}")
    torch.cuda.set_device(device)
    model, missing, unexpected = build_model(run_cfg, device)
    print(missing)
    print(unexpected)
    out_dir = None
    if ckpt_cfg["commit"]:
        out_dir = os.path.join(ckpt_cfg["ckpt_path"], "sample", run_cfg["debug"].get("sample_save_offset") or "")
    saved = refine_dataset(model, dataset, out_dir, batch_size=run_cfg["runtime"]["batch_size"], device=device,
                           commit=ckpt_cfg["commit"])
    for d in saved:
        _logger.info("sample %s", d["info"])
    return saved


if __name__ == "__main__":
    main()
