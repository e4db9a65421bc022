"""Developer probe: loop one kernel for ~2 s while sampling nvidia-smi power / clocks, to tell power-capped from structural limits.
Usage: python scripts/gpu_power_probe.py mainloop|cublas|attn"""
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
