#!/bin/bash
# GPU session C: 16-warp epilogue variant: kernel tests, main-loop-only timings, bench with per-class times.
mkdir -p gpurun_out
LOG=gpurun_out/r1c.log
: > $LOG
for t in test_linear1_fused test_linear2_gated test_whole_sequence_attention; do
  echo "######## pytest $t" >> $LOG
  timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k $t 2>&1 | tail -25 >> $LOG
done
echo "######## mainloop timings" >> $LOG
timeout 300 python scripts/gpu_time_kernels.py >> $LOG 2>&1
for pp in 0 2 4 6 8; do LAMSLIDE_ATTN_POLY=$pp timeout 120 python scripts/gpu_time_kernels.py attn >> $LOG 2>&1; done
echo "######## bench" >> $LOG
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
cat gpurun_out/bench_c.json >> $LOG; tail -5 gpurun_out/bench_c.err >> $LOG
echo "######## parity" >> $LOG
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_kernels.py 2>&1 | tail -8 >> $LOG
if [ -n "$WITH_NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$WITH_NCU" -s 40 -c 4 -f -o gpurun_out/prof_r1c \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile $BENCH_ARGS > gpurun_out/ncu_full_c.log 2>&1
tail -3 gpurun_out/ncu_full_c.log >> $LOG
fi
tail -120 $LOG
