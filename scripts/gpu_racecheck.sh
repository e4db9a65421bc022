#!/bin/bash
# compute-sanitizer racecheck on kernels whose shared-memory protocol is __syncwarp-based (hazards through TMA / mbarrier are outside the tool's model)
mkdir -p gpurun_out
LOG=gpurun_out/racecheck.log
: > $LOG
for sel in "short_strided_attention" "first_stage_linear and (640-32 or 130-42 or 1234-96)" "layernorm_epilogue and (700-32 or 1-96)"; do
  echo "######## $sel" >> $LOG
  timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 --print-limit 8 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "$sel" 2>&1 \
    | grep -E "passed|failed|RACECHECK SUMMARY|hazard|Race reported|=========     at" | head -14 >> $LOG
done
cat $LOG
