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


def trace(rows=128000, H=384, M=1536):
    """event timeline of CTA 0 (LAMSLIDE_FUSED_TRACE): cycles between the events of the epilogue warp and the two issuers"""
    lib = L.load()
    u = torch.randn(rows, H, device="cuda").to(torch.bfloat16)
    act = torch.randn(rows, H + M, device="cuda").to(torch.bfloat16)
    w1 = (torch.randn(3 * H + M, H, device="cuda") / math.sqrt(H)).to(torch.bfloat16)
    w2 = (torch.randn(H, H + M, device="cuda") / math.sqrt(H + M)).to(torch.bfloat16)
    b1 = torch.randn(3 * H + M, device="cuda") * 0.1
    b2 = torch.randn(H, device="cuda") * 0.1
    gate = torch.randn(rows // 2000 + 1, H, device="cuda")
    h = torch.zeros(rows, H, device="cuda")
    buf = torch.zeros(3 * 4096, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: L.check(lib.lamslide_debug_fused_mlp(u.data_ptr(), act.data_ptr(), w1.data_ptr(), w2.data_ptr(), b1.data_ptr(), b2.data_ptr(),
                                                        gate.data_ptr(), h.data_ptr(), rows, H, M, 2000, st))
    call()
    torch.cuda.synchronize()
    os.environ["LAMSLIDE_FUSED_TRACE"] = str(buf.data_ptr())
    call()
    torch.cuda.synchronize()
    del os.environ["LAMSLIDE_FUSED_TRACE"]
    t = buf.cpu().view(3, 4096)
    names = {0: {1: "wait acc1_full", 2: "got acc1_full", 3: "gelu done, wait g_empty", 4: "got g_empty", 5: "wait out_full", 6: "got out_full", 7: "drain done"},
             1: {1: "acc1_empty ok -> issue G1", 2: "G1 issued", 3: "attn_done ok"},
             2: {1: "wait out_free", 2: "out_free ok", 3: "attn phase issued", 4: "gg_full[0] ok", 5: "gg_full[1] ok", 6: "all issued"}}
    ev = []
    for r in range(3):
        n = int(t[r, 0])
        for i in range(n):
            x = int(t[r, 1 + i])
            ev.append((x & ((1 << 48) - 1), r, x >> 48))
    ev.sort()
    t0 = ev[0][0]
    role = ["EPI ", "ISS1", "ISS2"]
    last = {0: t0, 1: t0, 2: t0}
    lim = int(sys.argv[2]) if len(sys.argv) > 2 else 260
    for (c, r, tag) in ev[:lim]:
        print(f"{c - t0:9d}  (+{c - last[r]:6d})  {'          ' * r}{role[r]} {names[r].get(tag, tag)}")
        last[r] = c
    print("total cycles", ev[-1][0] - t0, "events", len(ev))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "trace":
    trace()
    sys.exit(0)

if __name__ == "__main__":
    rows, H, M = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (512, 384, 1536)
    for mode in ("attn_only", "mlp_only", "both"):
        run(rows, H, M, mode)
    run(rows, H, 128, "mlp_only")
    run(128 * 148 * 2, H, M, "both")
