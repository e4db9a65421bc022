#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sec.json 2> gpurun_out/bench_sec.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_sec.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print(d['kernel_time_shares'])
for k,v in d['secondary'].items():
    if k!='peptide_sweep': print(k, json.dumps(v))
    else:
        for r in v: print(' ', r['batch'], r['num_steps'], round(r['value'],1), round(r['tflops'],1))
PY
