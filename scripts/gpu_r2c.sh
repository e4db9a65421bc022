#!/bin/bash
# round 2, call C: tcgen05 attention — pipeline-only variant timing + ncu full capture
mkdir -p gpurun_out
LOG=gpurun_out/r2c.log
: > $LOG
timeout 300 python scripts/gpu_time_kernels.py attn_tc 2>&1 | grep "attention 4AA" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_atc2 \
  python scripts/gpu_time_kernels.py attn_tc_once > gpurun_out/ncu_atc2.log 2>&1
tail -3 gpurun_out/ncu_atc2.log >> $LOG
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "sample0" 2>&1 | tail -3 >> $LOG
tail -30 $LOG
