#!/bin/bash
# What each B200-native piece buys on the headline workload: the bench step with one A/B switch of a debug build at a time (same box).
mkdir -p gpurun_out
LAMSLIDE_DEBUG_KNOBS=1 python -c "import lam_slide_b200.build as b; b.build(force=True)" > /dev/null
OUT=gpurun_out/ablation.txt
: > $OUT
run() {
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-profile 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-42s %8.1f traj/s %8.2f ms/step  clocks %s' % ('$*', d['value'], d['ms_per_step'], d['clocks']['sm_mhz']))" >> $OUT
}
run BASE=1
run LAMSLIDE_NO_FUSED_LN=1
run LAMSLIDE_NO_FUSED_SPATIAL_ATTN=1
run LAMSLIDE_ATTN_NO_TC=1
run LAMSLIDE_FS_NO_TCGEN05=1
run LAMSLIDE_FS_NO_FOLD=1 LAMSLIDE_FS_NO_LN_EPILOGUE=1
run LAMSLIDE_NO_FUSED_MLP=1
run LAMSLIDE_LEGACY_GEMM=1
run BASE=2
python -c "import lam_slide_b200.build as b; b.build(force=True)" > /dev/null
cat $OUT
