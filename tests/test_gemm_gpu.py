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


@pytest.mark.parametrize("which,M,d,K1,N2", [
    (0, 300, 256, 256, 512), (0, 1000, 512, 512, 2048), (0, 10560, 512, 512, 2048), (1, 10560, 512, 2048, 1536),
    (2, 777, 512, 1024, 0), (1, 10432, 256, 1024, 768), (2, 10560, 512, 2048, 0), (0, 58, 512, 512, 1024),
    (1, 256, 512, 2048, 1536), (0, 19000, 512, 512, 2048)])
def test_chain_kernel_matches_fp32(which, M, d, K1, N2):
    """One launch of a chain kernel (csrc/gemm_chain.cuh): X = LayerNorm(X + A1 . W1^T + b1) in place as the bf16 hi/lo
    pair, then C2 = act(Xh . W2^T + b2), against torch fp32 on the same bf16-rounded inputs.  Covers one and two column
    halves (d = 256 / 512), row tiles that end inside the matrix, one unit per pair up to several rounds, no phase 2."""
    from tamf_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(M + d + K1 + N2)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    a1 = rn(M, K1).to(torch.bfloat16)
    w1 = (rn(d, K1) / K1 ** 0.5).to(torch.bfloat16)
    lnp = torch.cat([0.1 * rn(d), 1 + 0.1 * rn(d), 0.1 * rn(d)]).contiguous()
    x = rn(M, d)
    xh = x.to(torch.bfloat16)
    xl = (x - xh.float()).to(torch.bfloat16)
    x_in = xh.float() + xl.float()
    w2 = (rn(max(N2, 256), d) / d ** 0.5).to(torch.bfloat16)
    b2 = 0.1 * rn(max(N2, 256))
    c2 = torch.full((M, max(N2, 256)), float("nan"), device="cuda", dtype=torch.bfloat16)
    nb = L.tamf_chain_aux_bytes(M, d, max(K1, N2, 3 * d))
    aux = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    _lib.check(L.tamf_chain_run(which, _lib.ptr(a1), _lib.ptr(w1), _lib.ptr(lnp), _lib.ptr(xh), _lib.ptr(xl), _lib.ptr(w2),
                                _lib.ptr(b2), _lib.ptr(c2), M, d, K1, N2, _lib.ptr(aux), nb, None, _lib.stream_ptr()),
               "tamf_chain_run")
    torch.cuda.synchronize()
    y = x_in + a1.float() @ w1.float().t() + lnp[:d]
    ref = torch.nn.functional.layer_norm(y, (d,), lnp[d:2 * d], lnp[2 * d:], 1e-5)
    got = xh.float() + xl.float()
    err = (got - ref).abs().max().item()
    assert err < 5e-4, f"LayerNorm max abs err {err}"
    # the hi plane is bf16(x) (it doubles as the next GEMM's operand), the lo plane the rounding remainder: <= half an ulp
    assert bool((xl.float().abs() <= 2.0 ** -8 * xh.float().abs() + 1e-30).all())
    if which != 2:
        r2 = xh.float() @ w2[:N2].float().t() + b2[:N2]
        if which == 0:
            r2 = torch.nn.functional.gelu(r2)
        e2 = (c2[:, :N2].float() - r2).abs()
        assert bool((e2 <= 1e-2 + 1e-2 * r2.abs()).all()), f"phase-2 max abs err {e2.max().item()}"
