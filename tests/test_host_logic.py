"""Host-side logic of the drop-in classes that needs no GPU: schedule tables, state_dict contract, argument checks,
sharding, synthetic generators, bench bookkeeping."""
import numpy as np
import pytest
import torch

import tamf_b200
from tamf_b200 import shard, synth


def test_schedule_tables_match_reference_golden(golden):
    g = golden("diffusion_tables.npz")
    d = tamf_b200.create_gaussian_diffusion(1000, "cosine")
    assert d.num_timesteps == 1000
    for k in ("betas", "alphas_cumprod", "posterior_log_variance_clipped", "posterior_mean_coef1",
              "posterior_mean_coef2", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
        np.testing.assert_allclose(getattr(d, k), g[k], rtol=1e-12, atol=0)
    with pytest.raises(NotImplementedError):
        tamf_b200.create_gaussian_diffusion(1000, "cosine", sigma_small=False)
    with pytest.raises(NotImplementedError):
        tamf_b200.create_gaussian_diffusion(1000, "sqrt")


@pytest.mark.parametrize("arch,nparam", [("arch_mdm", 7031395), ("arch_mdm_l", 27300963)])
def test_state_dict_contract(arch, nparam):
    """Same parameter names / count as the reference without CLIP, strict=False load like
    launch/sample.py:190-196."""
    cfg = synth.ARCH[arch]
    m = tamf_b200.InterationSegmentMDM(**cfg, text_encoder=synth.text_features)
    sd = synth.g_state_dict(cfg, 0)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing and not unexpected
    assert sum(p.numel() for p in m.parameters_wo_clip()) == nparam
    keys = set(m.state_dict().keys())
    for k in ("seqTransEncoder.layers.7.self_attn.in_proj_weight", "input_merge.2.bias", "embed_text.weight",
              "embed_timestep.time_embed.0.weight", "output_process.poseFinal.bias", "hand_side_process.lh_embed",
              "sequence_pos_encoder.pe", "obj_input_process.poseEmbedding.weight"):
        assert k in keys
    extra = dict(sd)
    extra["clip_model.positional_embedding"] = torch.zeros(3)
    missing, unexpected = m.load_state_dict(extra, strict=False)
    assert unexpected == ["clip_model.positional_embedding"]


def test_hand_side_ids_and_errors():
    M = tamf_b200.InterationSegmentMDM
    assert M.hand_side_ids(["rh", "lh", "rh"]) == [0, 1, 0]
    with pytest.raises(ValueError, match="unexpected hand_side"):
        M.hand_side_ids(["rh", "both"])
    with pytest.raises(NotImplementedError):
        M(**{**synth.ARCH["arch_mdm"], "activation": "relu"})


def test_sampler_rejects_unsupported_hooks():
    d = tamf_b200.create_gaussian_diffusion(1000, "cosine")
    with pytest.raises(NotImplementedError):
        d.p_sample_loop(None, (1, 99, 1, 8), clip_denoised=True, model_kwargs={"batch": {}})
    with pytest.raises(NotImplementedError):
        d.p_sample_loop(None, (1, 99, 1, 8), clip_denoised=False, cond_fn=lambda *a: 0, model_kwargs={"batch": {}})


def test_q_sample_matches_closed_form():
    d = tamf_b200.create_gaussian_diffusion(1000, "cosine")
    x0, n = torch.ones(2, 3), torch.full((2, 3), 2.0)
    t = torch.tensor([0, 999])
    out = d.q_sample(x0, t, n)
    exp = d.sqrt_alphas_cumprod[[0, 999]][:, None] + 2 * d.sqrt_one_minus_alphas_cumprod[[0, 999]][:, None]
    np.testing.assert_allclose(out.numpy(), np.broadcast_to(exp, (2, 3)).astype(np.float32), rtol=1e-6)


def test_shard_ranges_cover_and_match_reference_rule():
    for n in (0, 1, 7, 64, 8192, 8191):
        for w in (1, 2, 3, 4, 8):
            rs = [shard.shard_range(n, r, w) for r in range(w)]
            assert rs[0].start == 0 and rs[-1].stop == n
            assert all(a.stop == b.start for a, b in zip(rs, rs[1:]))
            # launch/sample.py:198-199
            assert all(r.start == n * i // w and r.stop == n * (i + 1) // w for i, r in enumerate(rs))
    assert [len(b) for b in shard.batches(range(10, 150), 64)] == [64, 64, 12]
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def test_synth_is_deterministic_and_ragged_padding_is_zero():
    a, b = synth.make_batch(4, 16, nobj=3, seed=5, ragged=True), synth.make_batch(4, 16, nobj=3, seed=5, ragged=True)
    assert torch.equal(a["obj_traj"], b["obj_traj"]) and a["hand_side"] == ["rh", "lh", "rh", "lh"]
    for i in range(4):
        n = int(a["obj_num"][i])
        assert float(a["obj_traj"][i, n:].abs().sum()) == 0 and float(a["obj_embedding"][i, n:].abs().sum()) == 0
        assert len(a["obj_list"][i]) == n
    assert torch.equal(synth.text_features(["x", "y"]), synth.text_features(["x", "y"]))


def test_manolayer_host_surface():
    layer = tamf_b200.ManoLayer(side="left", assets=synth.mano_assets("left"))
    assert layer.th_faces.shape == (1538, 3) and layer.get_mano_closed_faces().shape == (1552, 3)
    assert layer.th_posedirs.shape == (778, 3, 135) and layer.th_J_regressor.shape == (16, 778)
    with pytest.raises(NotImplementedError):
        tamf_b200.ManoLayer(side="right", center_idx=9, assets=synth.mano_assets("right"))


def test_chamfer_argument_checks():
    with pytest.raises(ValueError):
        tamf_b200.nn_query(torch.zeros(4, 3), torch.zeros(1, 4, 3))
    with pytest.raises(ValueError):
        tamf_b200.ChamferDistance()(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3), point_reduction="max")


def test_bench_flop_model():
    import bench
    cfg = synth.ARCH["arch_mdm_l"]
    assert abs(bench.flops_per_seq_step(cfg) - 8.951e9) < 2e6  # 8 951 M
    for chain, n in ((True, 3 + 1 + 2 * 8 + 1), (False, 3 + 5 * 8 + 1)):
        cls = bench.kernel_classes(cfg, 64, chain=chain)
        assert len(cls) == n
        assert abs(sum(f for _, f in cls) / 64 - bench.flops_per_seq_step(cfg)) / 8.95e9 < 0.01


@pytest.mark.parametrize("M,d,ff,n_inp,slots", [(10560, 512, 2048, 1536, 74), (10560, 512, 2048, 0, 74), (10432, 256, 1024, 768, 74),
                                                (300, 512, 1024, 1536, 74), (58, 256, 512, 0, 74), (19000, 512, 2048, 1536, 66)])
def test_layer_schedule_is_complete_and_deadlock_free(M, d, ff, n_inp, slots):
    """The static schedule of the layer kernel (csrc/layer_chain.cuh build_layer_schedule, host code of the library, no GPU
    needed): every unit exactly once; the two column halves of a LayerNorm row tile on neighbouring pairs 2k / 2k + 1;
    and a replay in which a pair may only run its next unit once that unit's producers have finished completes (every
    pair's list is a subsequence of one global topological order)."""
    import ctypes as C
    from tamf_b200 import _lib
    L = _lib.lib()
    off = np.zeros(slots + 1, np.int32)
    units = np.zeros(1 << 16, np.int32)
    mk = C.c_double()
    pairs = L.tamf_layer_schedule(M, d, ff, n_inp, slots, off.ctypes.data, units.ctypes.data, len(units), C.byref(mk))
    assert 1 <= pairs <= slots and mk.value > 0
    T, H, n1, n3 = (M + 255) // 256, d // 256, ff // 256, n_inp // 256
    lists = [[(int(c) >> 28, (int(c) >> 8) & 0xFFFFF, int(c) & 255) for c in units[off[p]:off[p + 1]]] for p in range(pairs)]
    allu = [u for l in lists for u in l]
    expect = {(0, m, h) for m in range(T) for h in range(H)} | {(1, m, n) for m in range(T) for n in range(n1)} | \
             {(2, m, h) for m in range(T) for h in range(H)} | {(3, m, n) for m in range(T) for n in range(n3)}
    assert len(allu) == len(expect) and set(allu) == expect
    for p, l in enumerate(lists):
        for k, m, n in l:
            if k in (0, 2):
                assert p % H == n  # the half a pair's LayerNorm parameters are staged for
                if H == 2:
                    assert (k, m, n ^ 1) in lists[p ^ 1]
    # replay: done[kind][m] counts finished units; a unit may run when its producers have all finished
    pos = [0] * pairs
    done = {k: [0] * T for k in range(4)}
    need = {1: (0, H), 2: (1, n1), 3: (2, H)}
    left = len(allu)
    while left:
        progressed = False
        for p in range(pairs):
            while pos[p] < len(lists[p]):
                k, m, n = lists[p][pos[p]]
                if k in need and done[need[k][0]][m] < need[k][1]:
                    break
                if k in (0, 2) and H == 2:  # the partner half must be the partner pair's next unit, or already done
                    q = p ^ 1
                    ahead = lists[q][pos[q]:pos[q] + 1]
                    if (k, m, n ^ 1) not in ahead and (k, m, n ^ 1) not in lists[q][:pos[q]]:
                        break
                done[k][m] += 1
                pos[p] += 1
                left -= 1
                progressed = True
        assert progressed, "the schedule deadlocks"


def test_encode_text_clip_branch(monkeypatch):
    """InterationSegmentMDM.encode_text without a `text_encoder`: loads the `clip` package like load_and_freeze_clip
    (interaction_segment_mdm.py:84-97: clip.load(version, device='cpu', jit=False), convert_weights, eval), tokenises
    with context_length 22 + truncate and zero-pads the tokens to 77 columns before encode_text (:111-132)."""
    import sys
    import types

    import torch

    import tamf_b200
    calls = {}

    class FakeClipModel(torch.nn.Module):
        def encode_text(self, tokens):
            calls["tokens"] = tokens.clone()
            return tokens[:, :4].to(torch.float16) * 0.5

    def load(version, device="cuda", jit=True):
        calls["load"] = (version, device, jit)
        return FakeClipModel(), None

    def tokenize(texts, context_length=77, truncate=False):
        calls["tokenize"] = (list(texts), context_length, truncate)
        return torch.arange(len(texts) * context_length, dtype=torch.int32).reshape(len(texts), context_length) + 1

    fake = types.ModuleType("clip")
    fake.load, fake.tokenize = load, tokenize
    fake.model = types.SimpleNamespace(convert_weights=lambda m: calls.setdefault("converted", True))
    monkeypatch.setitem(sys.modules, "clip", fake)
    m = tamf_b200.InterationSegmentMDM(latent_dim=256, ff_size=1024, num_layers=8, num_heads=4)
    out = m.encode_text(["pick up the cup", "pour"])
    assert calls["load"] == ("ViT-B/32", "cpu", False) and calls["converted"]
    assert calls["tokenize"] == (["pick up the cup", "pour"], 22, True)
    tok = calls["tokens"]
    assert tok.shape == (2, 77) and bool((tok[:, 22:] == 0).all()) and bool((tok[:, :22] > 0).all())
    assert out.dtype == torch.float32 and out.shape == (2, 4) and float(out[0, 0]) == 0.5
    m.encode_text(["again"])
    assert calls["tokenize"][0] == ["again"]  # the tower is loaded once


def test_stack_schedule_is_complete_and_deadlock_free():
    """The schedule of the stack form (TAMF_CHAIN=2: all layers in one launch, csrc/layer_chain.cuh build_stack_schedule):
    every unit of every layer exactly once (the last layer without the next in_proj); LayerNorm halves on neighbouring
    pairs; and a replay with the attention CTAs as agents of their own (unit g = (layer * B + b) * H + h on CTA g mod A, in
    order) completes: every agent's list is a subsequence of one topological order."""
    import ctypes as C
    from tamf_b200 import _lib
    lib = _lib.lib()
    M, d, ff, L, slots, A, S, heads, B = 10560, 512, 2048, 8, 60, 28, 165, 4, 64
    off = np.zeros(slots + 1, np.int32)
    units = np.zeros(1 << 17, np.int32)
    mk = C.c_double()
    pairs = lib.tamf_stack_schedule(M, d, ff, L, slots, A, S, heads, 0.0, off.ctypes.data, units.ctypes.data, len(units),
                                    C.byref(mk))
    assert pairs == slots and mk.value > 0
    T, H, n1, n3 = (M + 255) // 256, d // 256, ff // 256, 3 * d // 256
    dec = lambda c: ((int(c) >> 24) & 15, (int(c) >> 28) & 3, (int(c) >> 8) & 0xFFFF, int(c) & 255)  # layer, kind, m, n
    lists = [[dec(c) for c in units[off[p]:off[p + 1]]] for p in range(pairs)]
    allu = [u for l in lists for u in l]
    expect = set()
    for l in range(L):
        expect |= {(l, 0, m, h) for m in range(T) for h in range(H)} | {(l, 1, m, n) for m in range(T) for n in range(n1)}
        expect |= {(l, 2, m, h) for m in range(T) for h in range(H)}
        if l + 1 < L:
            expect |= {(l, 3, m, n) for m in range(T) for n in range(n3)}
    assert len(allu) == len(expect) and set(allu) == expect
    for p, l in enumerate(lists):
        for lay, k, m, n in l:
            if k in (0, 2):
                assert p % H == n and (lay, k, m, n ^ 1) in lists[p ^ 1]
    # replay: GEMM pairs + attention CTAs
    att_lists = [[g for g in range(a, L * B * heads, A)] for a in range(A)]
    att_pos, pos = [0] * A, [0] * pairs
    done = {}  # (layer, kind, m) -> finished units
    att_done = {}  # (layer, b) -> finished heads
    cnt = lambda key: done.get(key, 0)
    left = len(allu) + L * B * heads
    while left:
        progressed = False
        for a in range(A):
            while att_pos[a] < len(att_lists[a]):
                g = att_lists[a][att_pos[a]]
                lay, b = g // (B * heads), (g // heads) % B
                if lay > 0 and any(cnt((lay - 1, 3, m)) < n3 for m in range(b * S // 256, (b * S + S - 1) // 256 + 1)):
                    break
                att_done[(lay, b)] = att_done.get((lay, b), 0) + 1
                att_pos[a] += 1
                left -= 1
                progressed = True
        for p in range(pairs):
            while pos[p] < len(lists[p]):
                lay, k, m, n = lists[p][pos[p]]
                if k == 0:
                    b0, b1 = m * 256 // S, min(B - 1, (m * 256 + 255) // S)
                    ok = all(att_done.get((lay, b), 0) == heads for b in range(b0, b1 + 1))
                    ok = ok and (lay == 0 or cnt((lay - 1, 2, m)) == H)
                elif k == 1:
                    ok = cnt((lay, 0, m)) == H
                elif k == 2:
                    ok = cnt((lay, 1, m)) == n1
                else:
                    ok = cnt((lay, 2, m)) == H
                if ok and k in (0, 2):  # the partner half must be the partner pair's next unit, or already done
                    q = p ^ 1
                    ok = (lay, k, m, n ^ 1) in lists[q][pos[q]:pos[q] + 1] or (lay, k, m, n ^ 1) in lists[q][:pos[q]]
                if not ok:
                    break
                done[(lay, k, m)] = cnt((lay, k, m)) + 1
                pos[p] += 1
                left -= 1
                progressed = True
        assert progressed, "the stack schedule deadlocks"
