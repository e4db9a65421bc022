#!/bin/bash
# first-stage kernels: isolated tests, parity, a short bench and the launch list of the first-stage kernels
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "first_stage_linear" 2>&1 | tail -15
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_fs.json 2> gpurun_out/bench_fs.err
tail -c 1500 gpurun_out/bench_fs.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fs_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-secondary > /dev/null 2>&1
echo done
