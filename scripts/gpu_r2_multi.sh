#!/bin/bash
# multi-GPU check (gpurun --gpus 2): the NCCL tests of the sharded entry points + the bench line at N = 2
mkdir -p gpurun_out
LOG=gpurun_out/multi.log
: > $LOG
nvidia-smi -L >> $LOG
timeout 900 python -m pytest tests/test_dist_nccl.py tests/test_gpu_parity.py -q -m gpu -k "nccl or two_devices or trajectory_files" --tb=short 2>&1 | grep -E "^E  |passed|failed|FAILED|Error|skipped" | cut -c1-260 | head -40 >> $LOG
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 \
  > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err >> $LOG
python - >> $LOG 2>&1 <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print("N=2 traj/s", round(d['value'],1), "ms/step", round(d['ms_per_step'],2), "e2e", round(d['e2e']['value'],1), d['config']['parallelism'], d['clocks'])
PY
tail -30 $LOG
