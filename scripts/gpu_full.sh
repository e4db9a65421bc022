#!/bin/bash
# everything the driver runs at round end, on one GPU: pytest -m gpu, smoke, bench (+ optional ncu, $WITH_NCU=1)
mkdir -p gpurun_out
LOG=gpurun_out/full.log
: > $LOG
echo "######## pytest -m gpu" >> $LOG
timeout 2400 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E  |passed|failed|FAILED|Error|skipped" | cut -c1-260 | head -60 >> $LOG
echo "######## smoke" >> $LOG
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 >> $LOG
echo "######## bench" >> $LOG
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -3 gpurun_out/bench_full.err >> $LOG
python - >> $LOG 2>&1 <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print("traj/s", round(d['value'],1), "ms/step", round(d['ms_per_step'],2), "e2e", round(d['e2e']['value'],1), d['clocks'], "launches", d['gpu_launches'])
ms=d['ms_per_step']
for k,v in d['kernel_time_shares'].items(): print(f"  {k:14s} {v*100:5.1f}%  {v*ms:6.2f} ms")
print(d['roofline']); print(d['whole_step']); print("cpu", d['cpu_baseline']); print("gpu eager", d['gpu_eager_baseline'])
for k,v in (d.get('secondary') or {}).items():
    if k != 'peptide_sweep': print(k, json.dumps(v))
PY
tail -70 $LOG
