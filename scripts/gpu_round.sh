#!/bin/bash
# One scripted GPU session: parity tests -> bench -> ncu launch list -> ncu full capture of the top kernels.
mkdir -p gpurun_out
LOG=gpurun_out/round.log
: > $LOG
echo "######## diag sample" >> $LOG
timeout 600 python scripts/gpu_diag.py sample >> $LOG 2>&1
echo "######## pytest -m gpu" >> $LOG
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 >> $LOG
echo "######## bench" >> $LOG
timeout 900 python bench.py $BENCH_ARGS > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json >> $LOG; tail -5 gpurun_out/bench.err >> $LOG
if [ -z "$SKIP_NCU" ]; then
echo "######## ncu launch list" >> $LOG
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log >> $LOG
wc -l gpurun_out/launches.csv >> $LOG
echo "######## ncu full: linear1 / linear2 / attention" >> $LOG
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_ws_kernel|mlp_fused_kernel|attn_seq_kernel|attn_rows_kernel|ln_modulate_kernel|linear_f32_v2_kernel' -s 120 -c 14 -f -o gpurun_out/prof_top \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log >> $LOG
ls -la gpurun_out >> $LOG
fi
tail -60 $LOG
