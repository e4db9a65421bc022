#!/bin/bash
# ncu --set full of the tcgen05 first-stage linear kernel on the [256000, 128] -> 128 layer (decoder ff / head layers)
mkdir -p gpurun_out
cat > /tmp/l5.py <<'PY'
import sys, math, torch
sys.path.insert(0, '.')
from lam_slide_b200 import _lib as L
lib = L.load()
rows, N, K = 256000, 128, 128
x = torch.randn(rows, K, device='cuda'); y = torch.empty(rows, N, device='cuda')
w = (torch.randn(N, K) / math.sqrt(K)).contiguous(); b = torch.randn(N).contiguous()
for _ in range(2):
    L.check(lib.lamslide_debug_fs_linear(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), rows, N, K, K, N, 0, 0, 0, 0, 1, 0, 0, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_tc5 -c 1 -s 1 -o gpurun_out/prof_l5 -f python /tmp/l5.py > gpurun_out/ncu_l5.log 2>&1
tail -5 gpurun_out/ncu_l5.log
