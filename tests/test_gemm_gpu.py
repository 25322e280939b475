"""tcgen05 GEMM self-test: C = A . W^T + b with bf16 inputs / fp32 accumulation vs torch fp32 matmul of the same
bf16-rounded inputs (tolerance = fp32 summation-order noise only)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("tile_n", [128, 256, 512])
@pytest.mark.parametrize("M,N,K", [(128, 512, 64), (128, 512, 512), (300, 1536, 512), (10560, 512, 2048),
                                   (58, 512, 128), (4096, 2048, 512)])
def test_gemm_matches_fp32(tile_n, cta_group, M, N, K):
    from tamf_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    c = torch.full((M, N), float("nan"), device="cuda")
    _lib.check(_lib.lib().tamf_gemm_selftest(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(c), M, N, K, tile_n,
                                             cta_group, _lib.stream_ptr()), "tamf_gemm_selftest")
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    err = (c - ref).abs().max().item()
    assert err < 2e-3, f"max abs err {err}"
