#!/bin/bash
# One scripted GPU session that regenerates the evidence under profiles/ from the commit it runs on:
#   parity tests -> bench line -> ncu launch list of one step -> ncu --set full of the top kernels -> traffic.json
# Usage (dev container): gpurun --timeout 3000 -- 'TAG=r02 bash scripts/gpu_round.sh'   then   python scripts/collect_profiles.py r02
TAG=${TAG:-r02}
mkdir -p gpurun_out
LOG=gpurun_out/round.log
: > $LOG
echo "######## pytest -m gpu" >> $LOG
timeout 2400 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E  |passed|failed|FAILED|Error|skipped" | cut -c1-260 | head -40 >> $LOG
echo "######## smoke" >> $LOG
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-200 >> $LOG
echo "######## bench" >> $LOG
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err >> $LOG
echo "######## ncu launch list (one step)" >> $LOG
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-secondary > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200 >> $LOG
wc -l gpurun_out/${TAG}_launches.csv >> $LOG
echo "######## ncu --set full: top kernels" >> $LOG
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'attn_tc_kernel|mlp_fused_kernel|gemm_ws_kernel|ln_modulate_kernel|linear_tc5_kernel' -s 150 -c 16 -f -o gpurun_out/${TAG}_top \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-secondary > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-200 >> $LOG
ls -la gpurun_out | grep ${TAG} >> $LOG
tail -40 $LOG
