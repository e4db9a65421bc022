"""Developer aid: the attention kernels on the MD17 spatial shape (S = 192 latents per frame, hd = 16)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lam_slide_b200 import _lib as L  # noqa: E402
from tests.test_gpu_kernels import _attention_reference  # noqa: E402

lib = L.load()
st = torch.cuda.current_stream().cuda_stream
mode = int(sys.argv[1])
B, T, Lx, H, heads = int(sys.argv[2]) if len(sys.argv) > 2 else 64, 30, 192, 256, 16
n = B * T * Lx
qkv = (torch.randn(n, 3 * H, device="cuda") * 0.6).to(torch.bfloat16)
out = torch.zeros(n, H, dtype=torch.bfloat16, device="cuda")
f = lambda: L.check(lib.lamslide_debug_attention(qkv.data_ptr(), out.data_ptr(), B, T, Lx, H, heads, H, 0, mode, st))
f()
torch.cuda.synchronize()
if B <= 4:
    ref = _attention_reference(qkv, B, T, Lx, H, heads, False)
    print(f"mode {mode} B={B}: max_rel {float((out.float() - ref).abs().max() / ref.abs().max()):.3e}", flush=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    f()
b.record()
torch.cuda.synchronize()
print(f"mode {mode} B={B}: {a.elapsed_time(b) * 100:.1f} us", flush=True)
