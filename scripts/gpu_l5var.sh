#!/bin/bash
# what bounds the tcgen05 first-stage linear kernel: CUDA-event times of one layer with parts of the kernel switched off (debug build)
mkdir -p gpurun_out
LAMSLIDE_DEBUG_KNOBS=1 python -c "import lam_slide_b200.build as b; b.build(force=True)" > /dev/null
cat > /tmp/l5v.py <<'PY'
import sys, os, math, torch
sys.path.insert(0, '.')
from lam_slide_b200 import _lib as L
lib = L.load()
def run(rows, N, K):
    x = torch.randn(rows, K, device='cuda'); y = torch.empty(rows, N, device='cuda')
    w = (torch.randn(N, K) / math.sqrt(K)).contiguous(); b = torch.randn(N).contiguous()
    st = torch.cuda.current_stream().cuda_stream
    f = lambda: L.check(lib.lamslide_debug_fs_linear(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), rows, N, K, K, N, 0, 0, 0, 0, 1, 0, 0, st))
    f(); f()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): f()
        torch.cuda.synchronize()
    ts = [e.device_time for e in prof.events() if 'linear_tc5' in e.name]
    return sum(ts) / max(len(ts), 1)
for shape in [(256000, 128, 128), (128000, 96, 96), (256000, 256, 256), (256000, 96, 384)]:
    print(shape, 'debug', os.environ.get('LAMSLIDE_L5_DEBUG', '0'), '%.1f us' % run(*shape), flush=True)
PY
for d in 0 1 2 4 8 12 14; do LAMSLIDE_L5_DEBUG=$d timeout 300 python /tmp/l5v.py 2>&1 | grep debug; done
python -c "import lam_slide_b200.build as b; b.build(force=True)" > /dev/null
