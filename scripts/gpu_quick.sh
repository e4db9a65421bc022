#!/bin/bash
# quick GPU check: selected kernel tests ($TESTS), bench, parity, optional ncu ($WITH_NCU = kernel regex)
mkdir -p gpurun_out
LOG=gpurun_out/quick.log
: > $LOG
for t in $TESTS; do
  echo "######## pytest $t" >> $LOG
  timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k $t 2>&1 | tail -15 >> $LOG
done
if [ -n "$MAINLOOP" ]; then timeout 300 python scripts/gpu_time_kernels.py >> $LOG 2>&1; fi
echo "######## bench" >> $LOG
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - >> $LOG 2>&1 <<'PY'
import json
d=json.load(open('gpurun_out/bench_q.json'))
print("traj/s", round(d['value'],1), "ms/step", round(d['ms_per_step'],2), "e2e", round(d['e2e']['value'],1), d['clocks'])
ms=d['ms_per_step']
for k,v in d['kernel_time_shares'].items(): print(f"  {k:14s} {v*100:5.1f}%  {v*ms:6.2f} ms")
print(d['roofline'])
PY
tail -3 gpurun_out/bench_q.err >> $LOG
echo "######## parity" >> $LOG
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_kernels.py 2>&1 | tail -6 >> $LOG
if [ -n "$WITH_NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$WITH_NCU" -s ${NCU_SKIP:-40} -c ${NCU_COUNT:-4} -f -o gpurun_out/prof_q \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile $BENCH_ARGS > gpurun_out/ncu_full_q.log 2>&1
tail -3 gpurun_out/ncu_full_q.log >> $LOG
fi
tail -100 $LOG
