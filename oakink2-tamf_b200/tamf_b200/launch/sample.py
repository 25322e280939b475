"""python -m tamf_b200.launch.sample -- sample every item of a split with MF-MDM G (ancestral DDPM, 1000 steps).

Entry point behind the reference's script/sample.sh:33-40:

    python -m tamf_b200.launch.sample --cfg config/obj_embedding.yml --cfg config/arch_mdm_l.yml \\
        --data.process_range "?(file:./asset/split/test.txt)" --data.cache_dict_filepath common/.../test.pkl \\
        --debug.model_weight_filepath model.pt --debug.sample_save_offset test/arch_mdm_l__0399 \\
        --runtime.device_id 0,1,2,3 --commit

Same flags and output layout as src/oakink2_tamf/launch/sample.py (reg_entry :54-127, worker :146-239, main :242-295):
worker w of W takes items [len*w/W, len*(w+1)/W) and writes <ckpt_path>/sample/<sample_save_offset>/%06d.npy ([T,99]
fp32) per item when --commit is given (dry run otherwise).  Differences, all in how the work is laid out on the GPUs:
one worker process per listed device by default (runtime.num_worker defaults to len(device_id); the reference's 8
workers over 4 GPUs exist to hide its launch-bound B=1 chains), and a worker runs chains of up to runtime.batch_size
sequences of equal frame and object count at once (extract_sample.sample_dataset) instead of B=1.  Under torchrun the
ranks are the workers and nothing is spawned."""
from __future__ import annotations

import logging
import os
from typing import List

import torch
import torch.multiprocessing as mp

from . import config as C
from .data import open_dataset

PROG = "sample"
WS_DIR = os.getcwd()
_logger = logging.getLogger(__name__)


def reg_entry(reg: C.Registry) -> None:
    reg.register("data_prefix", prefix="data", category=str, default=f"{WS_DIR}/data", abspath=True, required=True)
    reg.register("obj_embedding_prefix", prefix="data", category=str, default=None, abspath=True)
    reg.register("process_range", prefix="data", category=List[str], seq=":",
                 default=[f"?(file:{WS_DIR}/asset/split/test.txt)"])
    reg.register("cache_dict_filepath", prefix="data", category=str, abspath=True,
                 default=f"{WS_DIR}/common/save_cache_dict/main/cache/test.pkl")
    reg.register("source", prefix="data", category=str, default=None,
                 desc="reference | items:FILE.pkl | synthetic:N[:T[:K]] (launch/data.py)")
    C.reg_model_param(reg, "model")
    reg.register("model_weight_filepath", prefix="debug", category=str, abspath=True)
    reg.register("sample_save_offset", prefix="debug", category=str)
    reg.register("random_init_seed", prefix="debug", category=int, default=None,
                 desc="no checkpoint: random weights from this seed (dry runs)")
    reg.register("num_worker", prefix="runtime", category=int, default=None)
    reg.register("device_id", prefix="runtime", category=List[int], seq=",", default=[0, 1, 2, 3])
    reg.register("batch_size", prefix="runtime", category=int, default=64)
    reg.register("seed", prefix="runtime", category=int, default=None, desc="Philox seed of the chains (None: from torch)")
    reg.register("text_encoder", prefix="runtime", category=str, default="clip",
                 desc="clip (the reference's CLIP ViT-B/32 text tower) | synthetic (hash features, dry runs)")


def reg_extract(reg: C.Registry) -> dict:
    return {p: reg.select(p) for p in ("data", "debug", "model", "runtime")}


def build_model(run_cfg: dict, device: torch.device):
    """InterationSegmentMDM from model.* + weights from debug.model_weight_filepath, strict=False, a missing
    `clip_model.*` is expected (launch/sample.py:176-196)."""
    from .. import InterationSegmentMDM, synth
    mc = run_cfg["model"]
    enc = synth.text_features if run_cfg["runtime"].get("text_encoder") == "synthetic" else None
    model = InterationSegmentMDM(**{k: mc[k] for k in C.MODEL_DEFAULTS}, text_encoder=enc)
    path = run_cfg["debug"].get("model_weight_filepath")
    if path:
        state_dict = torch.load(path, map_location="cpu")
    elif run_cfg["debug"].get("random_init_seed") is not None:
        state_dict = synth.g_state_dict(dict(mc), seed=int(run_cfg["debug"]["random_init_seed"]))
    else:
        raise SystemExit("--debug.model_weight_filepath (or --debug.random_init_seed for a dry run) is required")
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    missing = [k for k in missing if not k.startswith("clip_model")]
    unexpected = [k for k in unexpected if not k.startswith("clip_model")]
    return model.to(device).eval(), missing, unexpected


def sample_worker(worker_id: int, num_worker: int, device_id: int, ckpt_cfg: dict, run_cfg: dict) -> int:
    logging.basicConfig(level=logging.INFO, format=f"%(asctime)s [{worker_id:02d}] %(name)s %(levelname)s %(message)s")
    from .. import create_gaussian_diffusion, sample_dataset
    _logger.info("worker_id: %02d  device_id: %d", worker_id, device_id)
    device = torch.device(f"cuda:{device_id}")
    torch.cuda.set_device(device)
    dataset = open_dataset(run_cfg["data"], enable_obj_model=True)
    model, missing, unexpected = build_model(run_cfg, device)
    if worker_id == 0:
        _logger.info("missing_keys: %s", missing)
        _logger.info("unexpected_keys: %s", unexpected)
    diffusion = create_gaussian_diffusion(diffusion_steps=1000, noise_schedule="cosine")
    lo, hi = len(dataset) * worker_id // num_worker, len(dataset) * (worker_id + 1) // num_worker
    _logger.info("%06d %06d", lo, hi)
    out_dir = None
    if ckpt_cfg["commit"]:
        out_dir = os.path.join(ckpt_cfg["ckpt_path"], "sample", run_cfg["debug"].get("sample_save_offset") or "")
    res = sample_dataset(model, diffusion, dataset, out_dir, worker_id=worker_id, num_worker=num_worker,
                         batch_size=run_cfg["runtime"]["batch_size"], device=device, commit=ckpt_cfg["commit"],
                         seed=run_cfg["runtime"].get("seed"))
    for sid in sorted(res):
        _logger.info("sample %06d", sid)
    return len(res)


def main(argv=None) -> None:
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s %(message)s")
    reg = C.Registry(PROG)
    C.reg_ckpt(reg, exp_id_default="main")  # launch/sample.py:56
    reg_entry(reg)
    reg.parse(argv)
    ckpt_cfg, run_cfg = C.ckpt_extract(reg), reg_extract(reg)
    rank = int(os.environ["RANK"]) if "RANK" in os.environ else None
    C.ckpt_setup(ckpt_cfg, _logger, rank)
    C.ckpt_opt(ckpt_cfg, rank, ckpt=ckpt_cfg, run=run_cfg)
    _logger.info("ckpt_cfg: %s", ckpt_cfg)
    _logger.info("run_cfg: %s", run_cfg)
    devices = run_cfg["runtime"]["device_id"]
    if rank is not None:  # torchrun: the ranks are the workers
        world = int(os.environ["WORLD_SIZE"])
        sample_worker(rank, world, devices[int(os.environ.get("LOCAL_RANK", rank)) % len(devices)], ckpt_cfg, run_cfg)
        return
    num_worker = run_cfg["runtime"]["num_worker"] or len(devices)
    if num_worker == 1:
        sample_worker(0, 1, devices[0], ckpt_cfg, run_cfg)
    else:
        ctx = mp.get_context("spawn")  # launch/sample.py:268
        procs = [ctx.Process(target=sample_worker, args=(w, num_worker, devices[w % len(devices)], ckpt_cfg, run_cfg))
                 for w in range(num_worker)]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        bad = [w for w, p in enumerate(procs) if p.exitcode != 0]
        if bad:
            raise SystemExit(f"workers {bad} failed")
    _logger.info("conclude parallel worker")


if __name__ == "__main__":
    main()
