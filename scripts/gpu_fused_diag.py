"""Developer aid (not a test): where does the fused MLP kernel differ from the fp64 reference?  Error per 128-row m-block and per
64-column group, with the attention half / the MLP half of W2 zeroed in turn.  Usage: python scripts/gpu_fused_diag.py [rows H M]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lam_slide_b200 import _lib as L  # noqa: E402


def run(rows, H, M, mode):
    lib = L.load()
    g = torch.Generator(device="cpu").manual_seed(rows + M + 5)
    u = torch.randn(rows, H, generator=g).to(torch.bfloat16).cuda()
    act = torch.randn(rows, H + M, generator=g).to(torch.bfloat16).cuda()
    w1 = (torch.randn(3 * H + M, H, generator=g) / math.sqrt(H)).to(torch.bfloat16).cuda()
    w2 = (torch.randn(H, H + M, generator=g) / math.sqrt(H + M)).to(torch.bfloat16).cuda()
    if mode == "attn_only":
        w2[:, H:] = 0
    if mode == "mlp_only":
        w2[:, :H] = 0
    b1 = (0.1 * torch.randn(3 * H + M, generator=g)).cuda()
    b2 = torch.zeros(H).cuda()
    gate = torch.ones(1, H).cuda()
    h = torch.zeros(rows, H).cuda()
    L.check(lib.lamslide_debug_fused_mlp(u.data_ptr(), act.data_ptr(), w1.data_ptr(), w2.data_ptr(), b1.data_ptr(), b2.data_ptr(),
                                         gate.data_ptr(), h.data_ptr(), rows, H, M, rows, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    pre = u.double() @ w1[3 * H:].double().t() + b1[3 * H:].double()
    hid = (0.5 * pre * (1.0 + torch.erf(pre / math.sqrt(2.0)))).to(torch.bfloat16).double()
    cat = torch.cat([act[:, :H].double(), hid], dim=1)
    ref = cat @ w2.double().t()
    err = (h.double() - ref).abs()
    scale = float(ref.abs().max())
    print(f"--- rows={rows} H={H} M={M} mode={mode}: max_rel {float(err.max()) / scale:.3e}")
    nmb = (rows + 127) // 128
    for mb in range(min(nmb, 8)):
        e = err[mb * 128:(mb + 1) * 128]
        cols = " ".join(f"{float(e[:, c:c + 64].max()) / scale:8.1e}" for c in range(0, H, 64))
        rws = " ".join(f"{float(e[r:r + 32].max()) / scale:8.1e}" for r in range(0, e.shape[0], 32))
        print(f"  m-block {mb}: by 64-col group [{cols}]   by 32-row group [{rws}]")


if __name__ == "__main__":
    rows, H, M = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (512, 384, 1536)
    for mode in ("attn_only", "mlp_only", "both"):
        run(rows, H, M, mode)
    run(rows, H, 128, "mlp_only")
    run(128 * 148 * 2, H, M, "both")
