"""Host logic of the G -> R pipeline module (tamf_b200/extract_sample.py): collate, selection, bihand view, launcher
sharding -- checked against the reference's own functions when /root/reference is present (this container)."""
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference/src"


def _items(**kw):
    from tamf_b200 import synth
    return synth.make_items(4, T=12, nobj=3, seed=3, npoints=32, **kw)


def test_collate_pads_object_axis_and_keeps_lists():
    from tamf_b200.extract_sample import interaction_segment_collate
    items = _items()
    b = interaction_segment_collate(items)
    nmax = max(it["obj_num"] for it in items)
    assert tuple(b["obj_traj"].shape) == (4, nmax, 12, 9) and tuple(b["obj_embedding"].shape) == (4, nmax, 768)
    for i, it in enumerate(items):
        k = it["obj_num"]
        assert np.array_equal(b["obj_traj"][i, :k].numpy(), it["obj_traj"])
        assert not b["obj_traj"][i, k:].any() and not b["obj_embedding"][i, k:].any()
    assert b["text"] == [it["text"] for it in items] and b["hand_side"] == [it["hand_side"] for it in items]
    assert b["obj_num"].dtype == torch.int64 and b["len"].tolist() == [12] * 4
    assert b["pose_repr"].dtype == torch.float32 and tuple(b["pose_repr"].shape) == (4, 12, 99)
    assert isinstance(b["obj_pointcloud"], list) and b["obj_pointcloud"][1] is items[1]["obj_pointcloud"]
    with pytest.raises(KeyError, match="unexpected key"):
        interaction_segment_collate([dict(items[0], bogus=1)])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_collate_and_select_match_reference():
    sys.path.insert(0, REF)
    try:
        from oakink2_tamf.dataset.collate import interaction_segment_collate as ref_collate
        from dev_fn.transform.cast import map_copy_select_to as ref_select
    finally:
        sys.path.remove(REF)
    from tamf_b200.extract_sample import SELECT_G, interaction_segment_collate, map_copy_select_to
    items = _items()
    a, r = interaction_segment_collate(items), ref_collate(items)
    assert list(a.keys()) == list(r.keys())
    for k in a:
        if isinstance(r[k], torch.Tensor):
            assert a[k].dtype == r[k].dtype and torch.equal(a[k], r[k]), k
        else:
            assert len(a[k]) == len(r[k]), k
    sa = map_copy_select_to(a, device="cpu", dtype=torch.float64, select=SELECT_G)
    sr = ref_select(r, device="cpu", dtype=torch.float64, select=SELECT_G)
    for k in sa:
        if isinstance(sr[k], torch.Tensor):
            assert sa[k].dtype == sr[k].dtype and torch.equal(sa[k], sr[k]), k
    assert sa["pose_repr"].dtype == torch.float64 and sa["sample_pose_repr"].dtype == torch.float32  # not selected


def test_bihand_item_selects_side_and_paired_objects():
    from tamf_b200.extract_sample import bihand_item
    it = _items(bihand=True)[2]
    rh, lh = bihand_item(it, "rh"), bihand_item(it, "lh")
    assert rh["pose_repr"] is it["pose_repr_rh"] and lh["pose_repr"] is it["pose_repr_lh"]
    assert rh["obj_list"] == it["obj_pair"][1] and lh["obj_list"] == it["obj_pair"][0]
    j = it["obj_list"].index(it["obj_pair"][1][0])
    assert np.array_equal(rh["obj_traj"][0], it["obj_traj"][j]) and rh["obj_num"] == 1
    assert np.array_equal(rh["obj_pointcloud"][0], it["obj_pointcloud"][j])
    with pytest.raises(ValueError, match="unexpected hand_side"):
        bihand_item(it, "xx")


def test_transf_merge_matches_oracle_transform():
    from oracle import tamf_oracle as orc
    from tamf_b200.extract_sample import transf_merge_obj_pointcloud
    it = _items()[3]
    got = transf_merge_obj_pointcloud(it["obj_pointcloud"], it["obj_traj"])
    k, T, P = it["obj_num"], 12, 32
    assert got.shape == (T, k * P, 3)
    for o in range(k):
        ref = orc.obj_world_points(torch.from_numpy(it["obj_traj"][o]), torch.from_numpy(it["obj_pointcloud"][o]))
        assert np.abs(got[:, o * P:(o + 1) * P] - ref.numpy()).max() < 1e-6


def test_same_len_batches_and_worker_shards_cover_the_dataset():
    from tamf_b200 import synth
    from tamf_b200.extract_sample import _same_len_batches
    from tamf_b200.shard import shard_range
    data = synth.make_items(5, T=8, npoints=8) + synth.make_items(3, T=10, npoints=8) + synth.make_items(2, T=8, npoints=8)
    seen = []
    for w in range(3):
        for grp in _same_len_batches(data, shard_range(len(data), w, 3), 2):
            assert len(grp) <= 2 and len({it["pose_repr"].shape[0] for _, it in grp}) == 1
            seen += [i for i, _ in grp]
    assert seen == list(range(10))


def test_launcher_groups_never_mix_object_counts():
    """A batched chain must be the stack of the reference's B = 1 chains: the collate zero-pads the object axis and the
    model averages over it, so items with different object counts (or frame counts) never share a group; one Philox
    stream per group."""
    from tamf_b200 import synth
    from tamf_b200.extract_sample import _group_seed, _same_len_batches, interaction_segment_collate
    items = synth.make_items(9, T=12, nobj=3, seed=5, npoints=16, ragged=True)
    counts = [it["obj_num"] for it in items]
    assert len(set(counts)) > 1  # the dataset really is ragged
    groups = list(_same_len_batches(items, range(len(items)), batch_size=4))
    assert [i for g in groups for i, _ in g] == list(range(9))  # order and coverage preserved
    for g in groups:
        assert len(g) <= 4 and len({it["obj_num"] for _, it in g}) == 1
        b = interaction_segment_collate([it for _, it in g])
        assert b["obj_traj"].shape[1] == g[0][1]["obj_num"]  # nothing was padded
    assert len({_group_seed(7, g) for g in groups}) == len(groups) and _group_seed(None, groups[0]) is None
    longer = synth.make_items(2, T=16, nobj=1, seed=1, npoints=16, ragged=False)
    mixed = [items[0], longer[0], longer[1]]
    assert [len(g) for g in _same_len_batches(mixed, range(3), batch_size=8)] == [1, 2]
