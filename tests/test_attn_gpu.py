"""Fused attention kernel vs a plain torch fp32 softmax(q k^T / sqrt(hd)) v of the same bf16-rounded inputs (the math
nn.TransformerEncoderLayer runs, no mask: interaction_segment_mdm.py:63-70,171).  Tolerance: P is rounded to bf16 before
P.V and the output is bf16, so 2 bf16 ulps of the output scale."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,S,H,d", [(2, 165, 4, 512), (3, 163, 4, 256), (1, 29, 4, 512), (2, 64, 4, 256),
                                     (1, 176, 4, 512), (5, 1, 4, 256), (64, 165, 4, 512)])
def test_attention_matches_fp32(B, S, H, d):
    from tamf_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + S)
    qkv = (1.5 * torch.randn(B * S, 3 * d, device="cuda", generator=g)).to(torch.bfloat16)
    out = torch.full((B * S, d), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().tamf_attn_selftest(_lib.ptr(qkv), _lib.ptr(out), B, S, H, d, _lib.stream_ptr()),
               "tamf_attn_selftest")
    torch.cuda.synchronize()
    hd = d // H
    q, k, v = (t.float().view(B, S, H, hd).transpose(1, 2) for t in qkv.split(d, dim=1))
    p = torch.softmax(q @ k.transpose(-1, -2) / hd ** 0.5, dim=-1)
    ref = (p @ v).transpose(1, 2).reshape(B * S, d)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().max().item()
    assert err < 3e-2 * max(1.0, ref.abs().max().item()), f"max abs err {err}"
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    assert rel < 8e-3, f"rel-L2 {rel}"
