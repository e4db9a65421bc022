#!/bin/bash
# tcgen05 attention: numerics probe + variant timings, kernel tests, optional ncu capture ($WITH_NCU=1)
mkdir -p gpurun_out
LOG=gpurun_out/attn.log
: > $LOG
timeout 300 python scripts/gpu_time_kernels.py attn_tc >> $LOG 2>&1
timeout 300 python scripts/gpu_time_kernels.py attn_tc_trace >> $LOG 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention" 2>&1 | tail -5 >> $LOG
if [ -n "$WITH_NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_atc3 \
  python scripts/gpu_time_kernels.py attn_tc_once > gpurun_out/ncu_atc3.log 2>&1
tail -3 gpurun_out/ncu_atc3.log >> $LOG
fi
tail -30 $LOG
