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


@pytest.mark.parametrize("M,d,ff,n_inp", [
    (300, 256, 512, 768), (1000, 512, 1024, 1536), (10560, 512, 2048, 1536), (10560, 512, 2048, 0), (777, 512, 1024, 0),
    (10432, 256, 1024, 768), (58, 512, 1024, 1536), (256, 512, 2048, 1536), (19000, 512, 2048, 1536)])
def test_layer_kernel_matches_fp32(M, d, ff, n_inp):
    """One launch of the layer kernel (csrc/layer_chain.cuh): X = LN1(X + ATT . Wo^T + b), H = gelu(Xh . W1^T + b),
    X = LN2(X + H . W2^T + b), QKV = Xh . Win^T + b -- against torch fp32 evaluated stage by stage on the kernel's own
    bf16 intermediates (so each stage is checked at fp32-summation tolerance).  Covers one and two column halves
    (d = 256 / 512), row tiles that end inside the matrix, fewer units than CTA pairs up to a dozen units per pair, and
    the last-layer form without the in_proj stage.  Run twice on the same scratch: the sync state must be reusable."""
    from tamf_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(M + d + ff + n_inp)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    bf = torch.bfloat16
    att = rn(M, d).to(bf)
    w_out, w1 = (rn(d, d) / d ** 0.5).to(bf), (rn(ff, d) / d ** 0.5).to(bf)
    w2, w_in = (rn(d, ff) / ff ** 0.5).to(bf), (rn(max(n_inp, 256), d) / d ** 0.5).to(bf)
    lnp = torch.cat([0.1 * rn(d), 1 + 0.1 * rn(d), 0.1 * rn(d), 0.1 * rn(d), 1 + 0.1 * rn(d), 0.1 * rn(d)]).contiguous()
    b1, b_in = 0.1 * rn(ff), 0.1 * rn(max(n_inp, 256))
    x = rn(M, d)
    nb = L.tamf_layer_aux_bytes(M, d, ff)
    aux = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    ln = torch.nn.functional.layer_norm
    for rep in range(2):
        xh = x.to(bf)
        xl = (x - xh.float()).to(bf)
        x_in = xh.float() + xl.float()
        H = torch.full((M, ff), float("nan"), device="cuda", dtype=bf)
        qkv = torch.full((M, max(n_inp, 256)), float("nan"), device="cuda", dtype=bf)
        _lib.check(L.tamf_layer_run(_lib.ptr(att), _lib.ptr(w_out), _lib.ptr(w1), _lib.ptr(w2), _lib.ptr(w_in), _lib.ptr(lnp),
                                    _lib.ptr(b1), _lib.ptr(b_in), _lib.ptr(xh), _lib.ptr(xl), _lib.ptr(H), _lib.ptr(qkv), M, d,
                                    ff, n_inp, _lib.ptr(aux), nb, None, _lib.stream_ptr()), "tamf_layer_run")
        torch.cuda.synchronize()
        # stage 1: LN1 (its output is only visible through H: check H against gelu(bf16(LN1) . W1^T + b1))
        y1 = ln(x_in + att.float() @ w_out.float().t() + lnp[:d], (d,), lnp[d:2 * d], lnp[2 * d:3 * d], 1e-5)
        h_ref = torch.nn.functional.gelu(y1.to(bf).float() @ w1.float().t() + b1)
        eh = (H.float() - h_ref).abs()
        # a bf16 rounding flip of a LN1 output moves an H element by ~2^-9 |w|: allow a few bf16 ulps
        assert bool((eh <= 3e-2 + 2e-2 * h_ref.abs()).all()), f"H max abs err {eh.max().item()}"
        # stage 3: LN2 from the kernel's own H and the fp32 LN1 output
        y2 = ln(y1 + H.float() @ w2.float().t() + lnp[3 * d:4 * d], (d,), lnp[4 * d:5 * d], lnp[5 * d:], 1e-5)
        got = xh.float() + xl.float()
        err = (got - y2).abs().max().item()
        assert err < 2e-3, f"LN2 max abs err {err}"  # y1 enters as the bf16 pair (2^-17 relative)
        assert bool((xl.float().abs() <= 2.0 ** -8 * xh.float().abs() + 1e-30).all())  # lo = rounding remainder of hi
        if n_inp:
            q_ref = xh.float() @ w_in[:n_inp].float().t() + b_in[:n_inp]
            eq = (qkv[:, :n_inp].float() - q_ref).abs()
            assert bool((eq <= 1e-2 + 1e-2 * q_ref.abs()).all()), f"QKV max abs err {eq.max().item()}"
