#!/bin/bash
# runs every diagnostic stage in its own process (a trapped kernel poisons only its stage), logs to gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/diag.log 2>&1
for stage in ${@:-gemm1 gemm attn first_stage backbone sample}; do
  echo "######## $stage" >> gpurun_out/diag.log
  timeout 300 python scripts/gpu_diag.py $stage >> gpurun_out/diag.log 2>&1
  echo "exit code $?" >> gpurun_out/diag.log
done
tail -120 gpurun_out/diag.log
