#!/bin/bash
# round 2, call D: fused spatial attention in linear1 + padded feature width: kernel tests, full parity, bench
mkdir -p gpurun_out
LOG=gpurun_out/r2d.log
: > $LOG
echo "######## linear1 kernel tests" >> $LOG
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "linear1" 2>&1 | tail -8 >> $LOG
echo "######## parity tests" >> $LOG
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_odeint.py -q -m gpu -x 2>&1 | tail -12 >> $LOG
echo "######## bench (1 GPU)" >> $LOG
timeout 1200 python bench.py --steps 5 --warmup 3 $BENCH_ARGS > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err
tail -5 gpurun_out/bench_r2d.err >> $LOG
python - >> $LOG 2>&1 <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2d.json'))
print("traj/s", round(d['value'],1), "ms/step", round(d['ms_per_step'],2), "e2e", round(d['e2e']['value'],1), d['clocks'], "launches", d['gpu_launches'])
ms=d['ms_per_step']
for k,v in d['kernel_time_shares'].items(): print(f"  {k:14s} {v*100:5.1f}%  {v*ms:6.2f} ms")
print(d['roofline'])
if d.get('secondary'):
    for k,v in d['secondary'].items():
        if k != 'peptide_sweep': print(k, json.dumps(v))
PY
tail -60 $LOG
