#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "fused_mlp or first_stage_linear" 2>&1 | tail -15
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_ln.json 2> gpurun_out/bench_ln.err
tail -c 1500 gpurun_out/bench_ln.json
echo done
