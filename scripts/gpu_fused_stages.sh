#!/bin/bash
# fused MLP kernel: sensitivity of the main loop to the ring depths (debug build), plain drain and LN drain
LAMSLIDE_DEBUG_KNOBS=1 python -c "import lam_slide_b200.build as b; b.build(force=True)" > /dev/null
cat > /tmp/fs.py <<'PY'
import os, sys, math, torch
sys.path.insert(0, '.')
from lam_slide_b200 import _lib as L
lib = L.load(); st = torch.cuda.current_stream().cuda_stream
rows, H, M = 128000, 384, 1536
u = torch.randn(rows, H, device="cuda").to(torch.bfloat16); act = torch.randn(rows, H + M, device="cuda").to(torch.bfloat16)
w1 = (torch.randn(3 * H + M, H, device="cuda") / math.sqrt(H)).to(torch.bfloat16); w2 = (torch.randn(H, H + M, device="cuda") / math.sqrt(H + M)).to(torch.bfloat16)
b1 = torch.randn(3 * H + M, device="cuda") * 0.1; b2 = torch.randn(H, device="cuda") * 0.1
nb = rows // 2000 + 1
gate = torch.randn(nb, H, device="cuda"); sh = torch.randn(nb, H, device="cuda"); sc = torch.randn(nb, H, device="cuda") * 0.3
h = torch.zeros(rows, H, device="cuda"); u2 = torch.empty_like(u)
def t(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / iters * 1e3
for stages in (8, 3, 2):
    os.environ["LAMSLIDE_FUSED_STAGES"] = str(stages)
    p = t(lambda: L.check(lib.lamslide_debug_fused_mlp(u.data_ptr(), act.data_ptr(), w1.data_ptr(), w2.data_ptr(), b1.data_ptr(), b2.data_ptr(), gate.data_ptr(), h.data_ptr(), rows, H, M, 2000, st)))
    l = t(lambda: L.check(lib.lamslide_debug_fused_mlp_ln(u.data_ptr(), act.data_ptr(), w1.data_ptr(), w2.data_ptr(), b1.data_ptr(), b2.data_ptr(), gate.data_ptr(), h.data_ptr(), rows, H, M, 2000, sh.data_ptr(), sc.data_ptr(), u2.data_ptr(), st)))
    print(f"ring depth cap {stages}: plain drain {p:.1f} us, LN drain {l:.1f} us", flush=True)
PY
timeout 300 python /tmp/fs.py 2>&1 | tail -4
python -c "import lam_slide_b200.build as b; b.build(force=True)" > /dev/null
