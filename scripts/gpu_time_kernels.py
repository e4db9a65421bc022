"""Developer timing aid (not a test): CUDA-event timings of the persistent GEMM main loop alone (no epilogue) for the
linear1 / linear2 shapes of the 4AA config, to separate "operand feed + tensor pipe" from "epilogue".
Usage: python scripts/gpu_time_kernels.py [rows]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lam_slide_b200 import _lib as L  # noqa: E402


def time_fn(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3  # us


def attn():
    """temporal attention of the 4AA config at B = 64 (128 sequences x 16 heads, S = 1000, hd = 24), whole-sequence kernel"""
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    B, T, Lx, H, heads = 64, 1000, 2, 384, 16
    n = B * T * Lx
    qkv = (torch.randn(n, 3 * H, device="cuda") * 0.6).to(torch.bfloat16)
    out = torch.zeros(n, H, dtype=torch.bfloat16, device="cuda")
    us = time_fn(lambda: L.check(lib.lamslide_debug_attention(qkv.data_ptr(), out.data_ptr(), B, T, Lx, H, heads, H, 1, 2, st)), iters=10)
    flops = 4.0 * 24 * T * T * heads * B * Lx
    print(f"attn_seq LAMSLIDE_ATTN_POLY={os.environ.get('LAMSLIDE_ATTN_POLY', 'default')}: {us:8.1f} us  {flops / us * 1e-6:7.1f} TFLOP/s", flush=True)


MODE = int(os.environ.get("ATC_MODE", "3"))


def attn_tc_probe():
    """numerics of the tcgen05 attention kernel on a few shapes (prints, does not assert) + timing of the exponential-mix variants
    (mode 3 + 4 v: v = 0 shipped mix, 1 all MUFU, 2 / 3: 2 / 4 of 8 pairs on the FMA-pipe polynomial) against the mma.sync kernel"""
    from tests.test_gpu_kernels import _attention_reference
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    for (B, T, Lx, H, heads) in [(1, 300, 2, 384, 16), (1, 512, 1, 256, 16), (2, 130, 3, 128, 4), (1, 64, 1, 256, 16), (2, 1000, 2, 384, 16)]:
        n = B * T * Lx
        g = torch.Generator().manual_seed(n)
        qkv = torch.randn(n, 3 * H, generator=g).to(torch.bfloat16).cuda()
        ref = _attention_reference(qkv, B, T, Lx, H, heads, True)
        out = torch.zeros(n, H, dtype=torch.bfloat16, device="cuda")
        rc = lib.lamslide_debug_attention(qkv.data_ptr(), out.data_ptr(), B, T, Lx, H, heads, H, 1, MODE, st)
        try:
            torch.cuda.synchronize()
            err = float((out.float() - ref).abs().max() / ref.abs().max())
            mean = float((out.float() - ref).abs().mean() / ref.abs().mean())
            print(f"attn_tc B={B} T={T} L={Lx} H={H} heads={heads}: rc={rc} max_rel={err:.3e} mean_rel={mean:.3e}", flush=True)
        except Exception as e:  # a trapped launch poisons the context: stop
            print(f"attn_tc: CUDA error {e}", flush=True)
            return
    B, T, Lx, H, heads = 64, 1000, 2, 384, 16
    n = B * T * Lx
    qkv = (torch.randn(n, 3 * H, device="cuda") * 0.6).to(torch.bfloat16)
    out = torch.zeros(n, H, dtype=torch.bfloat16, device="cuda")
    for mode, name in [(3, "tcgen05 shipped mix (3/8 poly)"), (7, "tcgen05 all MUFU"), (11, "tcgen05 2/8 poly"), (15, "tcgen05 4/8 poly"),
                       (19, "tcgen05 pipeline only (no exponentials)"),
                       (3 + 4 * 7, "tcgen05 3 groups, P in place (3/8 poly)"), (3 + 4 * 10, "tcgen05 3 groups (2/8 poly)"), (3 + 4 * 9, "tcgen05 3 groups (4/8 poly)"),
                       (3 + 4 * 8, "tcgen05 3 groups pipeline only"),
                       (2, "mma.sync whole-sequence")]:
        us = time_fn(lambda: L.check(lib.lamslide_debug_attention(qkv.data_ptr(), out.data_ptr(), B, T, Lx, H, heads, H, 1, mode, st)), iters=10)
        print(f"attention 4AA temporal [{name}]: {us:8.1f} us  {4.0 * 24 * T * T * heads * B * Lx / us * 1e-6:7.1f} TFLOP/s", flush=True)


def linear1():
    """linear1 of the 4AA config at B = 64 (128000 rows): full kernel, math without stores, stores without math"""
    import math
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 128000
    H, M, heads = 384, 1536, 16
    u = torch.randn(rows, H, device="cuda").to(torch.bfloat16)
    w1 = (torch.randn(3 * H + M, H, device="cuda") / math.sqrt(H)).to(torch.bfloat16)
    bias = torch.randn(3 * H + M, device="cuda") * 0.1
    gq = torch.ones(24, device="cuda")
    gk = torch.ones(24, device="cuda")
    qkv = torch.empty(rows, 3 * H, device="cuda", dtype=torch.bfloat16)
    act = torch.empty(rows, H + M, device="cuda", dtype=torch.bfloat16)
    for mode, name in [(32, "full, 2-CTA multicast"), (48, "full, 1 CTA"), (34, "math, no stores"), (35, "stores, no math"), (36, "no math, no stores"), (37, "full, evict-first stores"), (1, "legacy kernel")]:
        us = time_fn(lambda: L.check(lib.lamslide_debug_linear1(u.data_ptr(), w1.data_ptr(), bias.data_ptr(), gq.data_ptr(), gk.data_ptr(),
                                                                qkv.data_ptr(), act.data_ptr(), rows, H, M, heads, 2, 1000, 10000.0, mode, st)))
        print(f"linear1 rows={rows} [{name}]: {us:8.1f} us  {2.0 * rows * (3 * H + M) * H / us * 1e-6:7.1f} TFLOP/s", flush=True)


def fused():
    """fused MLP + linear2 kernel of the 4AA config at B = 64, with the profiling knobs of mlp_fused.cuh"""
    import math
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 128000
    H, M = 384, 1536
    u = torch.randn(rows, H, device="cuda").to(torch.bfloat16)
    act = torch.randn(rows, H + M, device="cuda").to(torch.bfloat16)
    w1 = (torch.randn(3 * H + M, H, device="cuda") / math.sqrt(H)).to(torch.bfloat16)
    w2 = (torch.randn(H, H + M, device="cuda") / math.sqrt(H + M)).to(torch.bfloat16)
    b1 = torch.randn(3 * H + M, device="cuda") * 0.1
    b2 = torch.randn(H, device="cuda") * 0.1
    gate = torch.randn(rows // 2000 + 1, H, device="cuda")
    h = torch.zeros(rows, H, device="cuda")
    flops = 2.0 * rows * (H * M + (H + M) * H)
    for dbg, stages in [(0, 8), (0, 4), (0, 3), (1, 8), (2, 8), (4, 8), (5, 8), (8, 8), (15, 8), (9, 8)]:
        os.environ["LAMSLIDE_FUSED_DEBUG"] = str(dbg)
        os.environ["LAMSLIDE_FUSED_STAGES"] = str(stages)
        us = time_fn(lambda: L.check(lib.lamslide_debug_fused_mlp(u.data_ptr(), act.data_ptr(), w1.data_ptr(), w2.data_ptr(), b1.data_ptr(),
                                                                  b2.data_ptr(), gate.data_ptr(), h.data_ptr(), rows, H, M, 2000, st)), iters=10)
        print(f"fused mlp rows={rows} debug={dbg} max_stages={stages}: {us:8.1f} us  {flops / us * 1e-6:7.1f} TFLOP/s", flush=True)


def attn_tc_once():
    """one warm launch pair of the shipped tcgen05 attention kernel at the 4AA shape (target of an ncu capture)"""
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    B, T, Lx, H, heads = 64, 1000, 2, 384, 16
    n = B * T * Lx
    qkv = (torch.randn(n, 3 * H, device="cuda") * 0.6).to(torch.bfloat16)
    out = torch.zeros(n, H, dtype=torch.bfloat16, device="cuda")
    for _ in range(3):
        L.check(lib.lamslide_debug_attention(qkv.data_ptr(), out.data_ptr(), B, T, Lx, H, heads, H, 1, 3, st))
    torch.cuda.synchronize()


def attn_tc_trace():
    """where the MMA warp and a softmax warp of the tcgen05 attention kernel spend their cycles (trace variants 5 / 6)"""
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    B, T, Lx, H, heads = 64, 1000, 2, 384, 16
    n = B * T * Lx
    qkv = (torch.randn(n, 3 * H, device="cuda") * 0.6).to(torch.bfloat16)
    out = torch.zeros(n, H, dtype=torch.bfloat16, device="cuda")
    buf = torch.zeros(148, 32, dtype=torch.int64, device="cuda")
    lib.lamslide_debug_attention_trace(buf.data_ptr())
    mma = ["wait s_free", "wait p_full", "issue QK (+q/kv waits)", "issue PV + commits", "-", "-", "-", "-"]
    smx = ["wait S + first ld", "-", "exponent phase", "-", "hand-off P", "epilogue", "-", "-"]
    for variant, name in [(5, "shipped mix"), (6, "no exponentials")]:
        buf.zero_()
        for _ in range(2):  # the kernel overwrites the buffer: the numbers are those of the second (warm) launch
            L.check(lib.lamslide_debug_attention(qkv.data_ptr(), out.data_ptr(), B, T, Lx, H, heads, H, 1, 3 + 4 * variant, st))
        torch.cuda.synchronize()
        t = buf.double().mean(0).cpu()
        steps = 13.84 * 32
        print(f"--- trace [{name}] cycles per 128-key chunk step (average over CTAs; {steps:.0f} steps per group and CTA)")
        print("  MMA warp 0   : " + ", ".join(f"{mma[i]} {t[i] / steps:.0f}" for i in range(0, 4)) + f"  | total {t[0:4].sum() / steps:.0f}")
        print("  softmax warp4: " + ", ".join(f"{smx[i]} {t[8 + i] / steps:.0f}" for i in range(6)) + f"  | total {t[8:16].sum() / steps:.0f}")
    lib.lamslide_debug_attention_trace(0)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "attn_tc_trace":
        return attn_tc_trace()
    if len(sys.argv) > 1 and sys.argv[1] == "attn_tc_once":
        return attn_tc_once()
    if len(sys.argv) > 1 and sys.argv[1] == "fused":
        return fused()
    if len(sys.argv) > 1 and sys.argv[1] == "attn":
        return attn()
    if len(sys.argv) > 1 and sys.argv[1] == "attn_tc":
        return attn_tc_probe()
    if len(sys.argv) > 1 and sys.argv[1] == "linear1":
        return linear1()
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 128000
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    for (N, K, bn) in [(2688, 384, 192), (2688, 384, -192), (384, 1920, 192), (384, 1920, -192), (384, 576, 192)]:
        a = torch.randn(rows, K, device="cuda").to(torch.bfloat16)
        b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        us = time_fn(lambda: L.check(lib.lamslide_debug_gemm_mainloop(a.data_ptr(), b.data_ptr(), rows, N, K, bn, st)))
        print(f"mainloop rows={rows} N={N} K={K} bn={bn}: {us:8.1f} us  {2.0 * rows * N * K / us * 1e-6:7.1f} TFLOP/s", flush=True)
        c = torch.empty(rows, N, device="cuda", dtype=torch.bfloat16)
        us = time_fn(lambda: torch.matmul(a, b.t(), out=c))
        print(f"   cuBLAS bf16 (same shape, bf16 out, no epilogue): {us:8.1f} us  {2.0 * rows * N * K / us * 1e-6:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
