"""The two launchers end to end on one GPU with synthetic items: python -m tamf_b200.launch.sample writes %06d.npy per item
(launch/sample.py:230-237), python -m tamf_b200.launch.sample_refine reads them back and writes save_dict.pkl per item
(launch/sample_refine.py:272-296); without --commit nothing is written."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run(mod, args, cwd):
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "oakink2-tamf_b200") + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", mod] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r


def test_sample_then_refine_launchers(tmp_path):
    from tamf_b200 import synth
    items = synth.make_items(5, T=24, nobj=2, seed=3, ragged=False)
    for it in items:
        it.pop("sample_pose_repr")
    with open(tmp_path / "items.pkl", "wb") as f:
        pickle.dump(items, f)
    common = ["--data.source", f"items:{tmp_path / 'items.pkl'}", "--debug.random_init_seed", "0"]
    g_args = common + ["--cfg", os.path.join(ROOT, "config", "arch_mdm.yml"), "--runtime.device_id", "0",
                       "--runtime.text_encoder", "synthetic", "--runtime.seed", "11", "--debug.sample_save_offset", "test/g"]
    _run("tamf_b200.launch.sample", g_args, tmp_path)
    assert not (tmp_path / "common").exists()  # dry run
    _run("tamf_b200.launch.sample", g_args + ["--commit"], tmp_path)
    out = tmp_path / "common" / "sample" / "main"
    files = sorted(os.listdir(out / "sample" / "test" / "g"))
    assert files == [f"{i:06d}.npy" for i in range(5)]
    a = np.load(out / "sample" / "test" / "g" / "000003.npy")
    assert a.shape == (24, 99) and a.dtype == np.float32 and np.isfinite(a).all()
    assert (out / "opt.yml").exists() and (out / "log.txt").exists()
    # the same seed gives the same samples (in-kernel Philox, one stream per group)
    _run("tamf_b200.launch.sample", g_args + ["--commit", "--exp_id", "again"], tmp_path)
    b = np.load(tmp_path / "common" / "sample" / "again" / "sample" / "test" / "g" / "000003.npy")
    assert np.array_equal(a, b)

    r_args = common + ["--cfg", os.path.join(ROOT, "config", "arch_refine.yml"), "--runtime.device_id", "0",
                       "--data.sample_dir", str(out / "sample" / "test" / "g"), "--debug.sample_save_offset", "test/r",
                       "--commit"]
    _run("tamf_b200.launch.sample_refine", r_args, tmp_path)
    rout = tmp_path / "common" / "sample_refine" / "main" / "sample" / "test" / "r"
    p = rout / "scene++003" / "3" / "0" / "save_dict.pkl"
    assert p.exists()
    with open(p, "rb") as f:
        d = pickle.load(f)
    assert set(d) == {"process_key", "info", "hand_side", "joints", "verts", "faces", "obj_list", "len", "frame_id",
                      "refine_pose_repr"}
    assert d["refine_pose_repr"].shape == (24, 99) and d["joints"].shape == (24, 21, 3) and d["verts"].shape == (24, 778, 3)
    assert d["process_key"] == "scene/003" and np.isfinite(d["verts"]).all()

    # contact-ratio score over what sample_refine wrote (script/compute_score/compute_score_cr.py)
    s_args = ["--data.source", f"items:{tmp_path / 'items.pkl'}", "--debug.synthetic_mano", "--runtime.device_id", "0",
              "--debug.sample_refine_filepath", str(rout)]
    r = _run("tamf_b200.launch.compute_score_cr", s_args, tmp_path)
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert lines[-2] == "(120,) (120,)"  # 5 items x 24 frames
    gt_ratio, refined_ratio = (float(v) for v in lines[-1].split())
    assert 0.0 <= gt_ratio <= 1.0 and 0.0 <= refined_ratio <= 1.0
    gt = np.load(tmp_path / "tmp" / "compute_score" / "contact_ratio" / "gt_contact_dist.npy")
    assert gt.shape == (120,) and np.isfinite(gt).all() and (gt >= 0).all()
    # the minimum over the nearest-neighbour distances is the minimum of the full distance matrix
    from tamf_b200 import transf_merge_obj_pointcloud
    it = items[0]
    merged = transf_merge_obj_pointcloud(np.asarray(it["obj_pointcloud"]), np.asarray(it["obj_traj"]))
    with open(rout / "scene++000" / "0" / "0" / "save_dict.pkl", "rb") as f:
        v0 = pickle.load(f)["verts"]
    ref = np.sqrt(((v0[:, :, None, :] - merged[:, None, :, :]) ** 2).sum(-1)).reshape(24, -1).min(1)
    got = np.load(tmp_path / "tmp" / "compute_score" / "contact_ratio" / "refined_contact_dist.npy")[:24]
    assert np.abs(got - ref).max() < 1e-5
