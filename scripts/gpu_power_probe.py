"""Developer probe: loop one kernel for ~2 s while sampling nvidia-smi power / clocks, to tell power-capped from structural limits.
Usage: python scripts/gpu_power_probe.py mainloop|mainloop1|cublas|bigcublas|attn|attn_tc|fused|linear1"""
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lam_slide_b200 import _lib as L  # noqa: E402


def main():
    what = sys.argv[1]
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    rows, N, K = 128000, 2688, 384
    a = torch.randn(rows, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    c = torch.empty(rows, N, device="cuda", dtype=torch.bfloat16)
    big = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
    bigc = torch.empty(8192, 8192, device="cuda", dtype=torch.bfloat16)
    if what == "mainloop":
        fn = lambda: L.check(lib.lamslide_debug_gemm_mainloop(a.data_ptr(), b.data_ptr(), rows, N, K, 192, st))
        flops = 2.0 * rows * N * K
    elif what == "mainloop1":
        fn = lambda: L.check(lib.lamslide_debug_gemm_mainloop(a.data_ptr(), b.data_ptr(), rows, N, K, -192, st))
        flops = 2.0 * rows * N * K
    elif what in ("attn", "attn_tc"):
        B, T, Lx, H, heads = 64, 1000, 2, 384, 16
        n = B * T * Lx
        qkv = (torch.randn(n, 3 * H, device="cuda") * 0.5).to(torch.bfloat16)
        out = torch.zeros(n, H, dtype=torch.bfloat16, device="cuda")
        mode = 2 if what == "attn" else 3
        fn = lambda: L.check(lib.lamslide_debug_attention(qkv.data_ptr(), out.data_ptr(), B, T, Lx, H, heads, H, 1, mode, st))
        flops = 4.0 * 24 * T * T * heads * B * Lx
    elif what == "fused":
        import math
        H, M = 384, 1536
        u = torch.randn(rows, H, device="cuda").to(torch.bfloat16)
        act = torch.randn(rows, H + M, device="cuda").to(torch.bfloat16)
        w1 = (torch.randn(3 * H + M, H, device="cuda") / math.sqrt(H)).to(torch.bfloat16)
        w2 = (torch.randn(H, H + M, device="cuda") / math.sqrt(H + M)).to(torch.bfloat16)
        b1 = torch.randn(3 * H + M, device="cuda") * 0.1
        b2 = torch.randn(H, device="cuda") * 0.1
        gate = torch.randn(rows // 2000 + 1, H, device="cuda")
        hh = torch.zeros(rows, H, device="cuda")
        fn = lambda: L.check(lib.lamslide_debug_fused_mlp(u.data_ptr(), act.data_ptr(), w1.data_ptr(), w2.data_ptr(), b1.data_ptr(), b2.data_ptr(),
                                                          gate.data_ptr(), hh.data_ptr(), rows, H, M, 2000, st))
        flops = 2.0 * rows * (H * M + (H + M) * H)
    elif what == "linear1":
        import math
        H, M, heads = 384, 0, 16
        u = torch.randn(rows, H, device="cuda").to(torch.bfloat16)
        w1 = (torch.randn(3 * H + 1536, H, device="cuda") / math.sqrt(H)).to(torch.bfloat16)
        bias = torch.randn(3 * H + 1536, device="cuda") * 0.1
        gq = torch.ones(24, device="cuda")
        gk = torch.ones(24, device="cuda")
        qkv = torch.empty(rows, 3 * H, device="cuda", dtype=torch.bfloat16)
        act = torch.empty(rows, H + 1536, device="cuda", dtype=torch.bfloat16)
        fn = lambda: L.check(lib.lamslide_debug_linear1(u.data_ptr(), w1.data_ptr(), bias.data_ptr(), gq.data_ptr(), gk.data_ptr(), qkv.data_ptr(),
                                                        act.data_ptr(), rows, H, 1536, heads, 2, 1000, 10000.0, 32, st))
        flops = 2.0 * rows * (3 * H + 1536) * H
    elif what == "cublas":
        fn = lambda: torch.matmul(a, b.t(), out=c)
        flops = 2.0 * rows * N * K
    else:
        fn = lambda: torch.matmul(big, big, out=bigc)
        flops = 2.0 * 8192 ** 3
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader", "-lms", "50"],
                           stdout=subprocess.PIPE, text=True)
    t0 = time.time()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 2.5:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    smi.terminate()
    lines = [l.strip() for l in smi.stdout.read().strip().splitlines() if l.strip()]
    mid = lines[len(lines) // 4: -max(1, len(lines) // 8)] or lines
    clocks = sorted(int(l.split(",")[0].split()[0]) for l in mid)
    power = sorted(float(l.split(",")[1].split()[0]) for l in mid)
    cap = sum("Active" in l.split(",")[2] and "Not" not in l.split(",")[2] for l in mid)
    print(f"{what} grid_cap={os.environ.get('LAMSLIDE_WS_GRID', '-')}: {ms / n * 1e3:8.1f} us/call {flops * n / ms * 1e-9:8.1f} TFLOP/s | "
          f"sm clock median {clocks[len(clocks) // 2]} MHz (min {clocks[0]}), power median {power[len(power) // 2]:.0f} W (max {power[-1]:.0f}), "
          f"sw_power_cap active in {cap}/{len(mid)} samples", flush=True)


if __name__ == "__main__":
    main()
