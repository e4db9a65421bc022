"""GPU: the tcgen05 GEMM and the attention kernels in isolation, through the C ABI test hooks."""
import math

import pytest
import torch

from tests.helpers import max_rel

pytestmark = pytest.mark.gpu


def _lib():
    from lam_slide_b200 import _lib as L
    return L


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 32, 64, 32), (128, 128, 128, 128), (300, 192, 384, 192), (1000, 96, 384, 96), (256, 256, 1920, 256),
    (4000, 384, 1920, 192), (77, 64, 256, 64), (515, 2688, 384, 192), (129, 48, 128, 48), (2048, 1280, 256, 128),
])
def test_tcgen05_gemm_matches_fp32_matmul(M, N, K, bn):
    L = _lib()
    lib = L.load()
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    a = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    b = (torch.randn(N, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    bias = torch.randn(N, generator=g).cuda()
    c = torch.full((M, N), float("nan"), device="cuda")
    L.check(lib.lamslide_debug_gemm(a.data_ptr(), b.data_ptr(), bias.data_ptr(), c.data_ptr(), M, N, K, bn,
                                    torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = a.double() @ b.double().t() + bias.double()
    assert torch.isfinite(c).all()
    assert max_rel(c, ref) < 2e-5, f"gemm {M}x{N}x{K} bn={bn}"


def _attention_reference(qkv, B, T, L, H, heads, temporal):
    hd = H // heads
    x = qkv.float().reshape(B, T, L, 3, heads, hd)
    q, k, v = x[:, :, :, 0], x[:, :, :, 1], x[:, :, :, 2]  # [B,T,L,h,hd]
    if temporal:
        q, k, v = (z.permute(0, 2, 3, 1, 4) for z in (q, k, v))  # [B,L,h,T,hd]
    else:
        q, k, v = (z.permute(0, 1, 3, 2, 4) for z in (q, k, v))  # [B,T,h,L,hd]
    s = (q @ k.transpose(-1, -2)) * math.log(2.0)  # q is pre-multiplied by hd^-0.5 * log2(e): softmax in base 2
    o = torch.softmax(s, dim=-1) @ v
    if temporal:
        o = o.permute(0, 3, 1, 2, 4)  # [B,T,L,h,hd]
    else:
        o = o.permute(0, 1, 3, 2, 4)
    return o.reshape(B * T * L, H)


@pytest.mark.parametrize("B,T,L,H,heads,temporal,flash", [
    (2, 1000, 2, 384, 16, 1, 0), (1, 300, 2, 384, 16, 1, 0), (2, 20, 8, 256, 16, 1, 0), (2, 20, 8, 256, 16, 1, 1),
    (3, 20, 2, 128, 4, 1, 1), (2, 7, 192, 256, 16, 0, 0), (4, 30, 2, 384, 16, 0, 0), (2, 20, 8, 256, 16, 0, 0),
    (2, 20, 8, 256, 16, 0, 1), (2, 130, 3, 128, 4, 1, 0), (1, 64, 1, 256, 16, 1, 0), (1, 129, 2, 384, 16, 1, 0),
])
def test_attention_matches_softmax_reference(B, T, L, H, heads, temporal, flash):
    L_ = _lib()
    lib = L_.load()
    n = B * T * L
    g = torch.Generator(device="cpu").manual_seed(n + H)
    qkv = (torch.randn(n, 3 * H, generator=g)).to(torch.bfloat16).cuda()
    ldo = H + 64
    out = torch.zeros(n, ldo, dtype=torch.bfloat16, device="cuda")
    L_.check(lib.lamslide_debug_attention(qkv.data_ptr(), out.data_ptr(), B, T, L, H, heads, ldo, temporal, flash,
                                          torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = _attention_reference(qkv, B, T, L, H, heads, bool(temporal))
    got = out[:, :H].float()
    assert torch.isfinite(got).all()
    assert float(out[:, H:].float().abs().max()) == 0.0  # columns beyond H untouched
    assert max_rel(got, ref) < 2e-2
    assert float((got - ref).abs().mean() / ref.abs().mean()) < 5e-3
