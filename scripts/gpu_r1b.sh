#!/bin/bash
# GPU session B (round 1): MUFU micro-benchmark, isolated kernel tests of the persistent GEMM / whole-sequence attention,
# bench, parity suite, ncu launch list + full capture of the new kernels.
mkdir -p gpurun_out
LOG=gpurun_out/r1b.log
: > $LOG
nvidia-smi -L >> $LOG
echo "######## mufu micro-benchmark" >> $LOG
(nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/mufu_bench.cu -o /tmp/mufu && timeout 120 /tmp/mufu) >> $LOG 2>&1
for t in test_whole_sequence_attention test_linear1_fused test_linear2_gated test_tcgen05_gemm test_attention_matches; do
  echo "######## pytest $t" >> $LOG
  timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k $t 2>&1 | tail -25 >> $LOG
done
echo "######## bench (new kernels)" >> $LOG
timeout 900 python bench.py --steps 3 --warmup 3 $BENCH_ARGS > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err
cat gpurun_out/bench_new.json >> $LOG; tail -5 gpurun_out/bench_new.err >> $LOG
echo "######## pytest -m gpu (parity)" >> $LOG
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_kernels.py 2>&1 | tail -15 >> $LOG
if [ -z "$SKIP_NCU" ]; then
echo "######## ncu launch list" >> $LOG
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile $BENCH_ARGS > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log >> $LOG
wc -l gpurun_out/launches.csv >> $LOG
echo "######## ncu full" >> $LOG
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_ws_kernel|attn_seq' -s 60 -c 6 -f -o gpurun_out/prof_r1b \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile $BENCH_ARGS > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log >> $LOG
fi
ls -la gpurun_out >> $LOG
tail -150 $LOG
