#!/bin/bash
mkdir -p gpurun_out
N=${N:-8}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 \
  > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print("N=$N traj/s", round(d['value'],1), "ms/step", round(d['ms_per_step'],2), "e2e", round(d['e2e']['value'],1), d['config']['parallelism'], d['clocks'])
PY
