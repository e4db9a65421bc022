#!/bin/bash
# round 2, call A: the new parity tests (untested configs, CUDA graphs, logit-bound fallback), then the reworked bench
mkdir -p gpurun_out
LOG=gpurun_out/r2a.log
: > $LOG
echo "######## new parity tests" >> $LOG
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "steps20 or steps50 or sample0 or large_batch or unbounded or cuda_graph or capturable or stream" 2>&1 | tail -25 >> $LOG
echo "######## bench (1 GPU, with secondary)" >> $LOG
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
tail -5 gpurun_out/bench_r2a.err >> $LOG
python - >> $LOG 2>&1 <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2a.json'))
print("traj/s", round(d['value'],1), "ms/step", round(d['ms_per_step'],2), "e2e", round(d['e2e']['value'],1), d['clocks'])
ms=d['ms_per_step']
for k,v in d['kernel_time_shares'].items(): print(f"  {k:14s} {v*100:5.1f}%  {v*ms:6.2f} ms")
print(d['roofline'])
print("cpu", d['cpu_baseline'])
print("gpu eager", d['gpu_eager_baseline'])
for k,v in d['secondary'].items(): print(k, json.dumps(v))
PY
tail -80 $LOG
