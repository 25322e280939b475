"""Launcher flag / config layer (tamf_b200/launch/config.py), the subset of config_reg + dev_fn.upkeep.ckpt the reference
launchers use (src/oakink2_tamf/launch/sample.py:54-143, sample_refine.py:48-128).  Host logic only: no GPU."""
import os
import pickle

import numpy as np
import pytest
import yaml

from conftest import ROOT


def _sample_registry(argv):
    from tamf_b200.launch import config as C
    from tamf_b200.launch import sample
    reg = C.Registry(sample.PROG)
    C.reg_ckpt(reg, exp_id_default="main")
    sample.reg_entry(reg)
    reg.parse(argv)
    return C.ckpt_extract(reg), sample.reg_extract(reg)


def test_sample_flags_like_script_sample_sh(tmp_path, monkeypatch):
    """The command line of script/sample.sh:33-40: two --cfg files, dotted overrides, comma list, --commit."""
    monkeypatch.chdir(tmp_path)
    split = tmp_path / "test.txt"
    split.write_text("scene_01\nscene_02\n\n")
    ckpt, run = _sample_registry([
        "--cfg", os.path.join(ROOT, "config", "obj_embedding.yml"), "--data.process_range", f"?(file:{split})",
        "--data.cache_dict_filepath", "common/save_cache_dict/main/cache/test.pkl",
        "--cfg", os.path.join(ROOT, "config", "arch_mdm_l.yml"), "--debug.model_weight_filepath", "w/model_0399.pt",
        "--debug.sample_save_offset", "test/arch_mdm_l__0399", "--runtime.device_id", "0,1,2,3", "--commit"])
    assert ckpt == {"exp_id": "main", "ckpt_path": str(tmp_path / "common" / "sample" / "main"),
                    "log_file": str(tmp_path / "common" / "sample" / "main" / "log.txt"), "commit": True}
    assert run["model"] == dict(input_dim=99, obj_input_dim=9, hand_shape_dim=10, obj_embed_dim=768, latent_dim=512,
                                ff_size=2048, num_layers=8, num_heads=4, dropout=0.1, activation="gelu")
    assert run["runtime"]["device_id"] == [0, 1, 2, 3] and run["runtime"]["batch_size"] == 64
    assert run["data"]["obj_embedding_prefix"] == str(tmp_path / "common/retrieve_obj_embedding/main/embedding")  # abspath
    assert run["debug"]["model_weight_filepath"] == str(tmp_path / "w/model_0399.pt")
    from tamf_b200.launch.config import expand_process_range
    assert expand_process_range(run["data"]["process_range"]) == ["scene_01", "scene_02"]


def test_defaults_precedence_and_dry_run(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    cfg = tmp_path / "a.yml"
    cfg.write_text(yaml.safe_dump({"model": {"latent_dim": 384, "num_layers": 6}, "runtime": {"device_id": [5, 6]}}))
    ckpt, run = _sample_registry(["--cfg", str(cfg), "--model.num_layers", "4", "--data.process_range", "a:b:c"])
    assert ckpt["commit"] is False  # dry run unless --commit (ckpt.py:108-120)
    assert run["model"]["latent_dim"] == 384 and run["model"]["num_layers"] == 4  # command line over config over default
    assert run["model"]["ff_size"] == 1024  # launch/param/model.py default
    assert run["runtime"]["device_id"] == [5, 6]
    assert run["data"]["process_range"] == ["a", "b", "c"]  # colon separated
    ckpt2, _ = _sample_registry(["--exp_id", "run_?(prog)"])
    assert ckpt2["exp_id"] == "run_sample" and ckpt2["ckpt_path"].endswith(os.path.join("common", "sample", "run_sample"))


def test_commit_writes_opt_yml_and_rotates(tmp_path, monkeypatch):
    import logging
    from tamf_b200.launch import config as C
    monkeypatch.chdir(tmp_path)
    ckpt, run = _sample_registry(["--commit", "--debug.sample_save_offset", "x"])
    lg = logging.getLogger("t")
    C.ckpt_setup(ckpt, lg)
    C.ckpt_opt(ckpt, ckpt=ckpt, run=run)
    C.ckpt_opt(ckpt, ckpt=ckpt, run=run)
    d = tmp_path / "common" / "sample" / "main"
    assert (d / "opt.yml").exists() and (d / "opt.yml.1").exists() and (d / "log.txt").exists()
    got = yaml.safe_load((d / "opt.yml").read_text())
    assert got["run"]["debug"]["sample_save_offset"] == "x" and got["ckpt"]["commit"] is True
    for h in list(logging.getLogger().handlers):
        if isinstance(h, logging.FileHandler):
            logging.getLogger().removeHandler(h)
            h.close()


def test_refine_flags_and_sample_adaptor(tmp_path, monkeypatch):
    from tamf_b200 import synth
    from tamf_b200.launch import config as C
    from tamf_b200.launch import sample_refine
    from tamf_b200.launch.data import GeneratedPoseReprSamples, open_dataset
    monkeypatch.chdir(tmp_path)
    reg = C.Registry(sample_refine.PROG)
    C.reg_ckpt(reg, exp_id_default="main")
    sample_refine.reg_entry(reg)
    items = synth.make_items(3, T=8, nobj=1, seed=1, ragged=False)
    with open(tmp_path / "items.pkl", "wb") as f:
        pickle.dump(items, f)
    reg.parse(["--data.source", f"items:{tmp_path / 'items.pkl'}", "--debug.random_init_seed", "0"])
    run = sample_refine.reg_extract(reg)
    assert run["runtime"]["device_id"] == 4  # launch/sample_refine.py:174
    assert run["data"]["sample_dir"] == ["common/sample/main/sample/test/arch_mdm_l__0399"]  # :172
    assert run["mano"]["mano_path"].endswith(os.path.join("asset", "mano_v1_2"))
    ds = open_dataset(run["data"], with_pointcloud=True)
    assert len(ds) == 3
    gen = tmp_path / "gen" / "arch_mdm_l__0399"
    gen.mkdir(parents=True)
    for i in range(3):
        np.save(gen / f"{i:06d}.npy", np.full((8, 99), float(i), np.float32))
    ad = GeneratedPoseReprSamples(ds, [str(gen)])
    it = ad[2]
    assert it["sample_info"] == ("arch_mdm_l__0399", 2) and float(it["sample_pose_repr"][0, 0]) == 2.0
    np.save(gen / "000003.npy", np.zeros((8, 99), np.float32))
    with pytest.raises(ValueError):
        GeneratedPoseReprSamples(ds, [str(gen)])  # one sample per item (pose_repr_sample.py:37)


def test_required_weights(tmp_path, monkeypatch):
    from tamf_b200.launch import sample
    monkeypatch.chdir(tmp_path)
    _, run = _sample_registry(["--data.source", "synthetic:2:8:1"])
    with pytest.raises(SystemExit):
        sample.build_model(run, "cpu")
