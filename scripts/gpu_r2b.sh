#!/bin/bash
# round 2, call B: the new tcgen05 attention kernel — numerics probe + timing, kernel tests, then parity + bench
mkdir -p gpurun_out
LOG=gpurun_out/r2b.log
: > $LOG
echo "######## attn_tc probe" >> $LOG
timeout 300 python scripts/gpu_time_kernels.py attn_tc >> $LOG 2>&1
echo "######## attention kernel tests" >> $LOG
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention" 2>&1 | tail -15 >> $LOG
echo "######## parity tests" >> $LOG
timeout 1800 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -25 >> $LOG
echo "######## bench (1 GPU, with secondary)" >> $LOG
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
tail -5 gpurun_out/bench_r2b.err >> $LOG
python - >> $LOG 2>&1 <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2b.json'))
print("traj/s", round(d['value'],1), "ms/step", round(d['ms_per_step'],2), "e2e", round(d['e2e']['value'],1), d['clocks'])
ms=d['ms_per_step']
for k,v in d['kernel_time_shares'].items(): print(f"  {k:14s} {v*100:5.1f}%  {v*ms:6.2f} ms")
print(d['roofline'])
print("cpu", d['cpu_baseline'])
print("gpu eager", d['gpu_eager_baseline'])
for k,v in d['secondary'].items(): print(k, json.dumps(v))
PY
tail -120 $LOG
