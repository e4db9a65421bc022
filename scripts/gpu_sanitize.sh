#!/bin/bash
# compute-sanitizer memcheck over the small-shape kernel tests (out-of-bounds / misaligned accesses in global and shared memory)
mkdir -p gpurun_out
LOG=gpurun_out/sanitize.log
: > $LOG
for sel in "first_stage_linear and (1-128 or 64-2 or 130-42 or 640-32 or 300-256 or 777-96 or 1234-96)" "layernorm_epilogue and (1-96 or 700-32 or 900-384)" \
           "short_strided_attention" "fused_mlp and (100-384 or 1-384)" "three_group"; do
  echo "######## $sel" >> $LOG
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "$sel" 2>&1 \
    | grep -E "passed|failed|ERROR SUMMARY|Invalid|invalid|Misaligned|out of bounds|=========     at" | head -12 >> $LOG
done
tail -60 $LOG
# ---- the whole path on the small goldens (first stage, conditioning, backbone, sample(), roll-out, SDE)
echo "######## parity: first stage / backbone / sample() on the small goldens" >> $LOG
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "first_stage_encode or setup_conditioning or backbone_forward or (sample_vs_reference_golden and (small or nba_full or pedestrian_full)) or rollout_vs_reference or sde_sampler or unbounded_logits" 2>&1 \
  | grep -E "passed|failed|ERROR SUMMARY|Invalid|invalid|Misaligned|out of bounds|=========     at" | head -12 >> $LOG
tail -8 $LOG
