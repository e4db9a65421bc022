#!/bin/bash
# quick check after a kernel change: the kernel tests selected by $K (default: all), the parity tests, a short bench with the kernel time shares
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu ${K:+-k "$K"} 2>&1 | tail -3
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); ms=d['ms_per_step']
print('traj/s %.1f  ms/step %.2f  e2e %.1f  clocks %s' % (d['value'], ms, d['e2e']['value'], d['clocks']['sm_mhz']))
print({k: round(v*ms,2) for k,v in d['kernel_time_shares'].items()})"
