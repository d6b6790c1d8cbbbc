#!/bin/bash
# One gpurun call that produces the raw material of profiles/r2_*:
#   gpurun --timeout 1500 -- 'bash tools/profile_round2.sh'
# 1. launch list (gpu__time_duration) of 3 graph-free config-A steps, 2. ncu --set full of the
# config-A step's kernels, 3. ncu --set full of one config-D (512x512, bf16, B=16) step.
set -u
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --ntime 600 --no-config-d --no-e2e-api"
$CMD > $OUT/r2_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/r2_plain.log; exit 1; }
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file $OUT/r2_launches_configA.csv $CMD > $OUT/r2_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:'conv3x3|dense|adam|sumsq|head_loss|build_batch|sgemm|skinny|outer' -s 110 -c 50 -o $OUT/r2_prof_configA -f \
  $CMD > $OUT/r2_ncu_full_A.log 2>&1; echo "config A full capture rc=$?"
DINCAE_NO_SIDE_STREAM=1 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'conv3x3|dense_tc|adam_factored|build_batch|gram|p4_to_p8|head_loss' -s 75 -c 75 -o $OUT/r2_prof_configD -f \
  python tools/configD_step.py 16 bf16 > $OUT/r2_ncu_full_D.log 2>&1; echo "config D full capture rc=$?"
ls -la $OUT/r2_prof_config*.ncu-rep
