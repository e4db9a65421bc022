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


@pytest.mark.parametrize("B,T,L,H,heads,temporal", [
    (2, 1000, 2, 384, 16, 1), (1, 300, 2, 384, 16, 1), (2, 7, 192, 256, 16, 0), (2, 130, 3, 128, 4, 1), (1, 129, 2, 384, 16, 1),
    (1, 64, 1, 256, 16, 1), (1, 1040, 1, 384, 16, 1), (3, 33, 2, 256, 16, 1),
])
def test_whole_sequence_attention_matches_softmax_reference(B, T, L, H, heads, temporal):
    """mode 2: K/V of the whole sequence resident in shared memory, softmax without the running maximum."""
    L_ = _lib()
    lib = L_.load()
    n = B * T * L
    g = torch.Generator(device="cpu").manual_seed(n + H + 1)
    qkv = (torch.randn(n, 3 * H, generator=g)).to(torch.bfloat16).cuda()
    ldo = H + 64
    out = torch.zeros(n, ldo, dtype=torch.bfloat16, device="cuda")
    L_.check(lib.lamslide_debug_attention(qkv.data_ptr(), out.data_ptr(), B, T, L, H, heads, ldo, temporal, 2,
                                          torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = _attention_reference(qkv, B, T, L, H, heads, bool(temporal))
    got = out[:, :H].float()
    assert torch.isfinite(got).all()
    assert float(out[:, H:].float().abs().max()) == 0.0
    assert max_rel(got, ref) < 2e-2
    assert float((got - ref).abs().mean() / ref.abs().mean()) < 5e-3


def _linear1_reference(u, w1, bias, gq, gk, H, M, heads, pos_div, pos_mod, theta=10000.0):
    hd = H // heads
    z = u.double() @ w1.double().t() + bias.double()
    rows = z.shape[0]
    qkv, mlp = z[:, :3 * H], z[:, 3 * H:]
    q, k, v = (qkv[:, i * H:(i + 1) * H].reshape(rows, heads, hd) for i in range(3))
    pos = (torch.arange(rows, device=z.device) // pos_div) % pos_mod
    omega = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.float64, device=z.device) / hd))
    ang = pos.double()[:, None] * omega[None, :]
    cos, sin = ang.cos()[:, None, :], ang.sin()[:, None, :]

    def norm_rope(x, g):
        x = x * torch.rsqrt((x * x).mean(-1, keepdim=True) + 1e-6) * g.double()
        xe, xo = x[..., 0::2], x[..., 1::2]
        return torch.stack([cos * xe - sin * xo, sin * xe + cos * xo], dim=-1).reshape(rows, heads, hd)

    q = norm_rope(q, gq) * (math.log2(math.e) / math.sqrt(hd))
    k = norm_rope(k, gk)
    qkv_ref = torch.cat([q.reshape(rows, H), k.reshape(rows, H), v.reshape(rows, H)], dim=1)
    mlp_ref = 0.5 * mlp * (1.0 + torch.erf(mlp / math.sqrt(2.0)))
    return qkv_ref, mlp_ref


@pytest.mark.parametrize("rows,H,M,heads,pos_div,pos_mod,legacy", [
    (1000, 384, 1536, 16, 2, 500, 0), (1000, 384, 1536, 16, 2, 500, 1), (20000 + 37, 384, 1536, 16, 1, 2, 0),
    (333, 256, 1024, 16, 8, 20, 0), (4096, 256, 512, 16, 192, 30, 0), (777, 128, 256, 4, 2, 20, 0), (128 * 150, 384, 1536, 16, 2, 1000, 0),
    (20000 + 37, 384, 1536, 16, 1, 2, 16), (128 * 151, 384, 1536, 16, 2, 1000, 0), (100, 384, 1536, 16, 2, 50, 0), (129, 256, 1024, 16, 8, 20, 0),
])
def test_linear1_fused_epilogue(rows, H, M, heads, pos_div, pos_mod, legacy):
    """linear1 + bias + QK-RMSNorm + RoPE + q pre-scale + erf-GELU (mmdit.py:241-247) vs an fp64 restatement."""
    L_ = _lib()
    lib = L_.load()
    hd = H // heads
    g = torch.Generator(device="cpu").manual_seed(rows + H)
    u = torch.randn(rows, H, generator=g).to(torch.bfloat16).cuda()
    w1 = (torch.randn(3 * H + M, H, generator=g) / math.sqrt(H)).to(torch.bfloat16).cuda()
    bias = (0.1 * torch.randn(3 * H + M, generator=g)).cuda()
    gq = (1.0 + 0.1 * torch.randn(hd, generator=g)).cuda()
    gk = (1.0 + 0.1 * torch.randn(hd, generator=g)).cuda()
    qkv = torch.full((rows, 3 * H), float("nan"), dtype=torch.bfloat16, device="cuda")
    act = torch.zeros(rows, H + M, dtype=torch.bfloat16, device="cuda")
    L_.check(lib.lamslide_debug_linear1(u.data_ptr(), w1.data_ptr(), bias.data_ptr(), gq.data_ptr(), gk.data_ptr(), qkv.data_ptr(),
                                        act.data_ptr(), rows, H, M, heads, pos_div, pos_mod, 10000.0, legacy,
                                        torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    qkv_ref, mlp_ref = _linear1_reference(u, w1, bias, gq, gk, H, M, heads, pos_div, pos_mod)
    assert torch.isfinite(qkv.float()).all()
    assert float(act[:, :H].float().abs().max()) == 0.0  # the attention half of act is not linear1's to write
    for name, got, ref in (("q", qkv[:, :H], qkv_ref[:, :H]), ("k", qkv[:, H:2 * H], qkv_ref[:, H:2 * H]),
                           ("v", qkv[:, 2 * H:], qkv_ref[:, 2 * H:]), ("mlp", act[:, H:], mlp_ref)):
        err = (got.double() - ref).abs()
        tol = 2.0 ** -8 * ref.abs() + 2e-3  # bf16 rounding of the result (+ small absolute slack)
        assert bool((err <= tol).all()), f"{name}: max err {float(err.max()):.3e} at {int(err.argmax())}"
        assert float(err.mean() / ref.abs().mean()) < 3e-3, name


@pytest.mark.parametrize("rows,H,M,heads,L", [
    (2000, 384, 1536, 16, 2), (20000 + 38, 384, 1536, 16, 2), (128 * 150, 384, 1536, 16, 2), (128 * 151 + 64, 384, 1536, 16, 2), (2, 384, 1536, 16, 2),
    (1600, 256, 1024, 16, 8), (8 * 2501, 256, 1024, 16, 8), (40 * 33, 128, 256, 4, 2), (4096 + 4, 384, 1536, 16, 4), (8 * 300, 384, 1536, 16, 8),
])
def test_linear1_with_fused_spatial_attention(rows, H, M, heads, L):
    """linear1 (q | k | v columns) + QK-RMSNorm + RoPE + the attention over sequences of L consecutive rows, all inside the GEMM epilogue
    (latent_si_v31.py:51-54; mmdit.py:42-55, 240-249), against an fp64 restatement: act[:, :H] = softmax(q k^T / sqrt(hd)) v."""
    L_ = _lib()
    lib = L_.load()
    hd = H // heads
    g = torch.Generator(device="cpu").manual_seed(rows + H + L)
    u = torch.randn(rows, H, generator=g).to(torch.bfloat16).cuda()
    w1 = (torch.randn(3 * H + M, H, generator=g) / math.sqrt(H)).to(torch.bfloat16).cuda()
    bias = (0.1 * torch.randn(3 * H + M, generator=g)).cuda()
    gq = (1.0 + 0.1 * torch.randn(hd, generator=g)).cuda()
    gk = (1.0 + 0.1 * torch.randn(hd, generator=g)).cuda()
    qkv = torch.zeros(rows, 3 * H, dtype=torch.bfloat16, device="cuda")
    act = torch.zeros(rows, H + M, dtype=torch.bfloat16, device="cuda")
    L_.check(lib.lamslide_debug_linear1(u.data_ptr(), w1.data_ptr(), bias.data_ptr(), gq.data_ptr(), gk.data_ptr(), qkv.data_ptr(),
                                        act.data_ptr(), rows, H, M, heads, 1, L, 10000.0, 64, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    qkv_ref, _ = _linear1_reference(u, w1, bias, gq, gk, H, M, heads, 1, L)
    q, k, v = (qkv_ref[:, i * H:(i + 1) * H].reshape(rows // L, L, heads, hd).permute(0, 2, 1, 3) for i in range(3))
    p = torch.softmax((q @ k.transpose(-1, -2)) * math.log(2.0), dim=-1)  # q carries hd^-0.5 * log2(e)
    ref = (p @ v).permute(0, 2, 1, 3).reshape(rows, H)
    got = act[:, :H].double()
    assert torch.isfinite(got).all()
    assert float(act[:, H:].float().abs().max()) == 0.0 and float(qkv.float().abs().max()) == 0.0  # nothing else is written
    err = (got - ref).abs()
    # q and k are rounded to bf16 before the logits (as in the unfused path, where they travel through HBM): ~2^-8 of a logit of
    # size ~5 moves a probability by up to ~1e-2
    assert bool((err <= 2.0 ** -6 * ref.abs() + 3e-2).all()), f"max err {float(err.max()):.3e} at {int(err.argmax())}"
    assert float(err.mean() / ref.abs().mean()) < 6e-3, f"mean rel err {float(err.mean() / ref.abs().mean()):.3e}"


@pytest.mark.parametrize("rows,H,M,rps,legacy", [
    (1000, 384, 1536, 250, 0), (1000, 384, 1536, 250, 1), (20000 + 37, 384, 1536, 2000, 0), (333, 256, 1024, 160, 0),
    (777, 128, 256, 40, 0), (128 * 150, 384, 1536, 2000, 0), (20000 + 37, 384, 1536, 2000, 16), (128 * 151, 384, 1536, 2000, 0),
    (100, 384, 1536, 50, 0),
])
def test_linear2_gated_residual(rows, H, M, rps, legacy):
    """h += gate[b] * (act @ w2^T + bias) (mmdit.py:248, latent_si_v31.py:54) vs fp64."""
    L_ = _lib()
    lib = L_.load()
    g = torch.Generator(device="cpu").manual_seed(rows + M)
    nb = (rows + rps - 1) // rps
    act = torch.randn(rows, H + M, generator=g).to(torch.bfloat16).cuda()
    w2 = (torch.randn(H, H + M, generator=g) / math.sqrt(H + M)).to(torch.bfloat16).cuda()
    bias = (0.1 * torch.randn(H, generator=g)).cuda()
    gate = torch.randn(nb, H, generator=g).cuda()
    h0 = torch.randn(rows, H, generator=g).cuda()
    h = h0.clone()
    L_.check(lib.lamslide_debug_linear2(act.data_ptr(), w2.data_ptr(), bias.data_ptr(), gate.data_ptr(), h.data_ptr(), rows, H, M,
                                        rps, legacy, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    b_of_row = torch.arange(rows, device="cuda") // rps
    ref = h0.double() + gate.double()[b_of_row] * (act.double() @ w2.double().t() + bias.double())
    assert torch.isfinite(h).all()
    assert max_rel(h, ref) < 2e-5


@pytest.mark.parametrize("B,T,L,H,heads,temporal,variant", [
    (2, 1000, 2, 384, 16, 1, 0), (1, 300, 2, 384, 16, 1, 0), (1, 129, 2, 384, 16, 1, 0), (1, 1024, 1, 384, 16, 1, 0), (2, 130, 3, 128, 4, 1, 0),
    (1, 512, 1, 256, 16, 1, 0), (2, 7, 192, 256, 16, 0, 0), (1, 64, 1, 256, 16, 1, 0), (6, 640, 2, 384, 16, 1, 0), (3, 1000, 2, 384, 16, 1, 0),
    (10, 100, 1, 384, 16, 1, 0), (2, 1000, 2, 384, 16, 1, 1), (2, 1000, 2, 384, 16, 1, 2), (2, 1000, 2, 384, 16, 1, 3), (5, 700, 2, 128, 4, 1, 0),
])
def test_tcgen05_attention_matches_softmax_reference(B, T, L, H, heads, temporal, variant):
    """mode 3 (+ 4 * variant: share of the exponentials on the FMA-pipe polynomial): the persistent tcgen05 kernel — S = Q K^T and
    O = P V on the 5th-gen tensor cores, accumulators and P in TMEM, K / V images in the no-swizzle UMMA layout, double buffered over
    (sequence, head) items.  Shapes: 1 .. 8 query tiles per item (odd and even counts: the two softmax groups interleave over the
    CTA's tile stream), fewer and more items than SMs, sequences that end inside a tile and inside a 32-key group, hd 16 / 24 / 32."""
    L_ = _lib()
    lib = L_.load()
    n = B * T * L
    g = torch.Generator(device="cpu").manual_seed(n + H + 2)
    qkv = (torch.randn(n, 3 * H, generator=g)).to(torch.bfloat16).cuda()
    ldo = H + 64
    out = torch.zeros(n, ldo, dtype=torch.bfloat16, device="cuda")
    L_.check(lib.lamslide_debug_attention(qkv.data_ptr(), out.data_ptr(), B, T, L, H, heads, ldo, temporal, 3 + 4 * variant,
                                          torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = _attention_reference(qkv, B, T, L, H, heads, bool(temporal))
    got = out[:, :H].float()
    assert torch.isfinite(got).all()
    assert float(out[:, H:].float().abs().max()) == 0.0
    assert max_rel(got, ref) < 2e-2
    assert float((got - ref).abs().mean() / ref.abs().mean()) < 5e-3
    # deterministic: same bits on a second launch (catches races between the softmax groups, the loaders and the MMA warp)
    out2 = torch.zeros_like(out)
    L_.check(lib.lamslide_debug_attention(qkv.data_ptr(), out2.data_ptr(), B, T, L, H, heads, ldo, temporal, 3 + 4 * variant,
                                          torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(out, out2)


@pytest.mark.parametrize("rows,H,M,rps", [
    (1000, 384, 1536, 250), (20000 + 37, 384, 1536, 2000), (128 * 150, 384, 1536, 2000), (333, 256, 1024, 160), (777, 128, 256, 40),
    (4096, 256, 512, 5760), (100, 384, 1536, 50), (1, 384, 1536, 7), (128 * 149 * 2 + 5, 256, 1024, 1000), (40000, 128, 128, 640),
    (256 * 75, 384, 128, 2000),
])
def test_fused_mlp_linear2(rows, H, M, rps):
    """h += gate[b] * ([attn | gelu(u W1m^T + b1m)] W2^T + b2): the MLP half of linear1, GELU and linear2 in one kernel
    (mmdit.py:241-248, latent_si_v31.py:54) vs fp64 (with the hidden activation rounded to bf16 as the A operand)."""
    L_ = _lib()
    lib = L_.load()
    g = torch.Generator(device="cpu").manual_seed(rows + M + 5)
    nb = (rows + rps - 1) // rps
    u = torch.randn(rows, H, generator=g).to(torch.bfloat16).cuda()
    act = torch.randn(rows, H + M, generator=g).to(torch.bfloat16).cuda()  # only [:, :H] (attention output) is read
    w1 = (torch.randn(3 * H + M, H, generator=g) / math.sqrt(H)).to(torch.bfloat16).cuda()
    w2 = (torch.randn(H, H + M, generator=g) / math.sqrt(H + M)).to(torch.bfloat16).cuda()
    b1 = (0.1 * torch.randn(3 * H + M, generator=g)).cuda()
    b2 = (0.1 * torch.randn(H, generator=g)).cuda()
    gate = torch.randn(nb, H, generator=g).cuda()
    h0 = torch.randn(rows, H, generator=g).cuda()
    h = h0.clone()
    L_.check(lib.lamslide_debug_fused_mlp(u.data_ptr(), act.data_ptr(), w1.data_ptr(), w2.data_ptr(), b1.data_ptr(), b2.data_ptr(),
                                          gate.data_ptr(), h.data_ptr(), rows, H, M, rps, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    pre = u.double() @ w1[3 * H:].double().t() + b1[3 * H:].double()
    hid = (0.5 * pre * (1.0 + torch.erf(pre / math.sqrt(2.0)))).to(torch.bfloat16).double()
    cat = torch.cat([act[:, :H].double(), hid], dim=1)
    b_of_row = torch.arange(rows, device="cuda") // rps
    ref = h0.double() + gate.double()[b_of_row] * (cat @ w2.double().t() + b2.double())
    assert torch.isfinite(h).all()
    assert max_rel(h, ref) < 2e-3  # bf16 rounding flips of the hidden activation (GELU fit 2.6e-5) only
    assert float((h.double() - ref).abs().mean() / ref.abs().mean()) < 2e-4


@pytest.mark.parametrize("rows,N,K,ldx,ldy,act,res,rowadd", [
    (256000 // 8, 256, 108, 108, 256, 1, False, 0),   # net_merge.0 (peptide): K tail of 12 inside the last 32-wide k-block
    (4000, 256, 256, 256, 384, 0, False, 4),          # net_merge.2 written into the [x | E_ent] context matrix + sin/cos row add
    (5000, 96, 384, 384, 96, 1, False, 0),            # encoder.mlp.0
    (5000, 384, 96, 96, 384, 0, False, 0),            # encoder.mlp.2: two 192-wide tiles
    (3000, 64, 384, 384, 64, 0, False, 0),            # to_kv
    (1234, 96, 32, 32, 96, 0, True, 0),               # to_out with the in-place residual
    (777, 96, 96, 96, 288, 0, False, 0),              # to_qkv-like pitch
    (2000, 768, 96, 96, 768, 0, False, 0),            # extender (4 x 192)
    (130, 42, 128, 128, 42, 0, False, 0),             # atom14_pos head: N % 16 != 0, scalar stores (pitch 42)
    (513, 20, 128, 128, 20, 0, False, 0),             # aatype head
    (64, 2, 128, 128, 2, 0, False, 0),                # 2-wide position head
    (1, 128, 128, 128, 128, 2, True, 0),              # a single row, SiLU of the sum
    (19000, 128, 128, 128, 128, 1, False, 0),         # more tiles than SMs: both accumulators and every ring phase
    (300, 256, 12, 12, 256, 0, False, 0),             # K smaller than one k-block
    (640, 32, 32, 32, 32, 0, False, 0),
])
@pytest.mark.parametrize("path", [0, 1])
def test_first_stage_linear_is_fp32_accurate(rows, N, K, ldx, ldy, act, res, rowadd, path):
    """One first-stage layer (torch_modules.py / encoder.py / decoder.py nn.Linear + its fused epilogue) against fp64: the
    tcgen05 3xTF32 kernel (path 0) and the mma.sync / FMA kernels (path 1) must both be as good as an fp32 FMA chain."""
    L_ = _lib()
    lib = L_.load()
    g = torch.Generator(device="cpu").manual_seed(rows * 3 + N + K)
    x = torch.randn(rows, ldx, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).contiguous()
    b = (0.3 * torch.randn(N, generator=g)).contiguous()
    y0 = torch.randn(rows, ldy, generator=g).cuda()
    y = y0.clone()
    ra = torch.randn(max(rowadd, 1), N, generator=g).cuda()
    L_.check(lib.lamslide_debug_fs_linear(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), rows, N, K, ldx, ldy, act,
                                          y.data_ptr() if res else 0, ldy, ra.data_ptr() if rowadd else 0, max(rowadd, 1), N, path,
                                          torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = x[:, :K].double() @ w.double().cuda().t() + b.double().cuda()
    if act == 1:
        ref = 0.5 * ref * (1.0 + torch.erf(ref / math.sqrt(2.0)))
    if rowadd:
        ref = ref + ra.double()[torch.arange(rows, device="cuda") % rowadd]
    if res:
        ref = ref + y0[:, :N].double()
    if act == 2:
        ref = ref * torch.sigmoid(ref)
    assert torch.isfinite(y).all()
    assert max_rel(y[:, :N], ref) < 5e-6, f"linear {rows}x{N}x{K} path={path}"
    if ldy > N:
        assert torch.equal(y[:, N:], y0[:, N:]), "columns past N must not be touched"


@pytest.mark.parametrize("rows,H,M,rps", [
    (4096, 384, 1536, 2000), (128 * 149 * 2 + 5, 384, 1536, 2000), (100, 384, 1536, 50), (1, 384, 1536, 7), (4096, 256, 512, 5760),
    (128 * 149 * 2 + 5, 256, 1024, 1000), (40000, 128, 128, 640), (256 * 75, 384, 128, 2000),
])
def test_fused_mlp_drain_writes_next_layernorm_modulate(rows, H, M, rps):
    """The fused MLP kernel in LN mode: besides h += gate * (...), its drain writes u' = LayerNorm(h_new) * (1 + scale) + shift
    (the next block's pre_norm + modulate, latent_si_v31.py:50,57) over the buffer its own A operand came from."""
    L_ = _lib()
    lib = L_.load()
    g = torch.Generator(device="cpu").manual_seed(rows + M + 11)
    nb = (rows + rps - 1) // rps
    u0 = torch.randn(rows, H, generator=g).to(torch.bfloat16).cuda()
    u = u0.clone()
    act = torch.randn(rows, H + M, generator=g).to(torch.bfloat16).cuda()
    w1 = (torch.randn(3 * H + M, H, generator=g) / math.sqrt(H)).to(torch.bfloat16).cuda()
    w2 = (torch.randn(H, H + M, generator=g) / math.sqrt(H + M)).to(torch.bfloat16).cuda()
    b1 = (0.1 * torch.randn(3 * H + M, generator=g)).cuda()
    b2 = (0.1 * torch.randn(H, generator=g)).cuda()
    gate = torch.randn(nb, H, generator=g).cuda()
    shift = torch.randn(nb, H, generator=g).cuda()
    scale = (0.5 * torch.randn(nb, H, generator=g)).cuda()
    # a residual stream with a row mean far from zero and very different row scales: the one-pass statistics must cope
    h0 = (torch.randn(rows, H, generator=g) * (0.1 + 10.0 * torch.rand(rows, 1, generator=g)) + 5.0 * torch.randn(rows, 1, generator=g)).cuda()
    h = h0.clone()
    st = torch.cuda.current_stream().cuda_stream
    L_.check(lib.lamslide_debug_fused_mlp_ln(u.data_ptr(), act.data_ptr(), w1.data_ptr(), w2.data_ptr(), b1.data_ptr(), b2.data_ptr(),
                                             gate.data_ptr(), h.data_ptr(), rows, H, M, rps, shift.data_ptr(), scale.data_ptr(),
                                             u.data_ptr(), st))
    torch.cuda.synchronize()
    # the residual update must equal the plain kernel's (TMA reduce-add path) to fp32 rounding
    h_plain = h0.clone()
    L_.check(lib.lamslide_debug_fused_mlp(u0.data_ptr(), act.data_ptr(), w1.data_ptr(), w2.data_ptr(), b1.data_ptr(), b2.data_ptr(),
                                          gate.data_ptr(), h_plain.data_ptr(), rows, H, M, rps, st))
    torch.cuda.synchronize()
    assert torch.isfinite(h).all() and torch.isfinite(u.float()).all()
    assert max_rel(h, h_plain.double()) < 1e-6
    # u' against fp64 LayerNorm + modulate of the h the kernel produced
    b_of_row = torch.arange(rows, device="cuda") // rps
    hd = h.double()
    ln = (hd - hd.mean(1, keepdim=True)) / torch.sqrt(hd.var(1, unbiased=False, keepdim=True) + 1e-6)
    ref = ln * (1.0 + scale.double()[b_of_row]) + shift.double()[b_of_row]
    err = (u.double() - ref).abs()
    assert float((err / (ref.abs() + 1.0)).max()) < 2 ** -8  # one bf16 rounding of the output
    assert float(err.mean() / ref.abs().mean()) < 2e-3


@pytest.mark.parametrize("rows,N,K,group,affine,res,y_needed", [
    (5000, 96, 96, 0, True, True, 1),      # to_out + residual -> ff.norm (row LayerNorm, the two column halves merge their statistics)
    (3001, 128, 32, 0, True, True, 1),     # decoder output block
    (700, 32, 32, 0, True, False, 1),      # D = 32 configs
    (4000, 96, 96, 0, False, False, 0),    # quant: Linear -> LayerNorm without affine, only the normalised output is needed
    (2000, 768, 96, 96, True, False, 0),   # extender: LayerNorm per 96-wide token of the row (norm_context of the output block)
    (900, 384, 96, 0, True, False, 1),     # a row wider than one tile: the LayerNorm runs as its own launch
    (1, 96, 96, 0, True, True, 1),
    (19000, 128, 128, 0, True, True, 1),   # more tiles than SMs
])
def test_first_stage_linear_with_layernorm_epilogue(rows, N, K, group, affine, res, y_needed):
    """A first-stage layer whose result feeds a LayerNorm (torch_modules.py PreNorm of the next sub-layer, lightning_base.py quant,
    decoder.py extender -> norm_context): y and LN(y) against fp64."""
    L_ = _lib()
    lib = L_.load()
    g = torch.Generator(device="cpu").manual_seed(rows + N + K + group)
    x = torch.randn(rows, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).contiguous()
    b = (0.3 * torch.randn(N, generator=g) + 0.5).contiguous()
    gw = max(group, 1) if group else N
    lw = (1.0 + 0.2 * torch.randn(gw, generator=g)).cuda()
    lb = (0.2 * torch.randn(gw, generator=g)).cuda()
    r0 = (3.0 * torch.randn(rows, N, generator=g) + 2.0).cuda()
    y = r0.clone()
    out = torch.full((rows, N), float("nan"), device="cuda")
    L_.check(lib.lamslide_debug_fs_linear_ln(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), out.data_ptr(),
                                             lw.data_ptr() if affine else 0, lb.data_ptr() if affine else 0, group, rows, N, K,
                                             y.data_ptr() if res else 0, y_needed, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = x.double() @ w.double().cuda().t() + b.double().cuda()
    if res:
        ref = ref + r0.double()
    if y_needed:
        assert max_rel(y, ref) < 5e-6
    t = ref.reshape(rows, N // gw, gw)
    ln = (t - t.mean(-1, keepdim=True)) / torch.sqrt(t.var(-1, unbiased=False, keepdim=True) + 1e-5)
    if affine:
        ln = ln * lw.double() + lb.double()
    ln = ln.reshape(rows, N)
    assert torch.isfinite(out).all()
    assert float((out.double() - ln).abs().max()) < 2e-5


def test_three_group_attention_variant_matches_reference():
    """attn_tc3.cuh (three tile groups, probabilities written in place over the logits): a measured alternative that is not the
    default; its numerics are pinned so the record of the experiment stays runnable."""
    L_ = _lib()
    lib = L_.load()
    for (B, T, L, H, heads) in [(2, 1000, 2, 384, 16), (1, 512, 1, 256, 16), (40, 300, 3, 128, 4), (1, 300, 2, 384, 16)]:
        n = B * T * L
        g = torch.Generator(device="cpu").manual_seed(n + 17)
        qkv = torch.randn(n, 3 * H, generator=g).to(torch.bfloat16).cuda()
        out = torch.zeros(n, H, dtype=torch.bfloat16, device="cuda")
        L_.check(lib.lamslide_debug_attention(qkv.data_ptr(), out.data_ptr(), B, T, L, H, heads, H, 1, 3 + 4 * 7, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        ref = _attention_reference(qkv, B, T, L, H, heads, True)
        assert max_rel(out.float(), ref) < 2e-2
        assert float((out.float() - ref).abs().mean() / ref.abs().mean()) < 5e-3


@pytest.mark.parametrize("B,T,L,H,heads", [
    (2, 20, 8, 256, 16),     # NBA temporal: hd 16, groups of 4 heads
    (3, 20, 2, 128, 4),      # pedestrian temporal: hd 32
    (2, 30, 192, 256, 16),   # MD17 temporal: token stride 192
    (2, 20, 8, 384, 16),     # hd 24
    (2, 25, 5, 96, 6),       # head count not a multiple of 4: groups of 2
    (1, 32, 3, 256, 16), (5, 1, 4, 256, 16), (1, 2, 1, 128, 4),
])
def test_short_strided_attention(B, T, L, H, heads):
    """Temporal attention of the small-T configurations (sequence over T, tokens L rows apart): the warp-per-(sequence, head group)
    kernel against the softmax reference, and the launch must be that kernel."""
    L_ = _lib()
    lib = L_.load()
    n = B * T * L
    g = torch.Generator(device="cpu").manual_seed(n + H + 3)
    qkv = (torch.randn(n, 3 * H, generator=g)).to(torch.bfloat16).cuda()
    ldo = H + 64
    out = torch.zeros(n, ldo, dtype=torch.bfloat16, device="cuda")
    L_.kernel_count("attn_short", reset=True)
    L_.check(lib.lamslide_debug_attention(qkv.data_ptr(), out.data_ptr(), B, T, L, H, heads, ldo, 1, 0, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    if L > 1:
        assert L_.kernel_count("attn_short") == 1
    ref = _attention_reference(qkv, B, T, L, H, heads, True)
    got = out[:, :H].float()
    assert torch.isfinite(got).all()
    assert float(out[:, H:].float().abs().max()) == 0.0
    assert max_rel(got, ref) < 2e-2
    assert float((got - ref).abs().mean() / ref.abs().mean()) < 5e-3
