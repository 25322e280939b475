"""Where the launchers take their items from.

The reference builds `InteractionSegmentData` over the OakInk2 release (dataset/interaction_segment.py:285-451), which
is data loading, outside the sampling hot path (out of scope).  The launchers therefore accept, through
`--data.source`:

  reference            the reference's own dataset class, constructed with the reference's arguments.  Needs the
                       `oakink2_tamf` package importable in the user's environment (default when it is).
  items:FILE.pkl       a pickled list of items in the schema of InteractionSegmentData.__getitem__ (:415-448)
                       -- what `pickle.dump([ds[i] for i in range(len(ds))], f)` writes in the reference's environment.
  synthetic:N[:T[:K]]  N seeded synthetic items of T frames with K objects (dry runs, CI).
"""
from __future__ import annotations

import os
import pickle
from typing import List, Sequence

import numpy as np

from .config import expand_process_range


class ItemList:
    """A list of items with the two attributes the launchers read from the reference dataset."""

    def __init__(self, items: Sequence[dict], obj_store=None):
        self.items, self.obj_store = list(items), obj_store

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return dict(self.items[i])


def default_source() -> str:
    try:
        import oakink2_tamf.dataset.interaction_segment  # noqa: F401
        return "reference"
    except Exception:
        return ""


def open_dataset(data_cfg: dict, enable_obj_model: bool = True, with_pointcloud: bool = False):
    src = data_cfg.get("source") or default_source()
    if not src:
        raise SystemExit("--data.source is required here: the reference package `oakink2_tamf` is not importable "
                         "(use items:FILE.pkl or synthetic:N)")
    if src == "reference":
        from oakink2_tamf.dataset.interaction_segment import InteractionSegmentData
        with open(data_cfg["cache_dict_filepath"], "rb") as f:
            cache_dict = pickle.load(f)
        kw = dict(process_range_list=expand_process_range(data_cfg["process_range"]), data_prefix=data_cfg["data_prefix"],
                  obj_embedding_prefix=data_cfg["obj_embedding_prefix"], enable_obj_model=enable_obj_model,
                  cache_dict=cache_dict)
        if with_pointcloud:  # launch/sample_refine.py:161-169
            kw.update(obj_pointcloud_prefix=data_cfg["obj_pointcloud_prefix"], append_reverse_segment=False)
        return InteractionSegmentData(**kw)
    if src.startswith("items:"):
        with open(src[6:], "rb") as f:
            return ItemList(pickle.load(f))
    if src.startswith("synthetic:"):
        from .. import synth
        parts = [int(p) for p in src[10:].split(":")]
        n, T, k = parts[0], (parts[1] if len(parts) > 1 else 160), (parts[2] if len(parts) > 2 else 2)
        return ItemList(synth.make_items(n, T=T, nobj=k, seed=0, ragged=False))
    raise SystemExit(f"unknown --data.source {src!r}")


class GeneratedPoseReprSamples:
    """Pairs item i of the dataset with the i-th generated sample `<dir>/%06d.npy` (what launch/sample.py saved):
    adds `sample_info` and `sample_pose_repr` (dataset/pose_repr_sample.py:18-52)."""

    def __init__(self, dataset, dir_list: List[str]):
        self.dataset, self.info, self.samples = dataset, [], {}
        for d in dir_list:
            base = os.path.basename(os.path.normpath(d))
            for fn in sorted(f for f in os.listdir(d) if os.path.splitext(f)[-1] == ".npy"):
                key = (base, int(os.path.splitext(fn)[0]))
                self.info.append(key)
                self.samples[key] = np.load(os.path.join(d, fn))
        if len(self.info) != len(dataset):
            raise ValueError(f"{len(self.info)} generated samples for {len(dataset)} dataset items")
        self.obj_store = getattr(dataset, "obj_store", None)

    def __len__(self):
        return len(self.info)

    def __getitem__(self, i):
        item = self.dataset[i]
        item["sample_info"], item["sample_pose_repr"] = self.info[i], self.samples[self.info[i]]
        return item
