#!/bin/bash
# One gpurun call that produces the raw material of profiles/r2_*; the .ncu-rep files stay on the
# box (they exceed gpurun's 64 MiB return limit), only CSV extracts come back:
#   gpurun --timeout 1500 -- 'bash tools/profile_round2.sh'
set -u
OUT=gpurun_out
TMP=/tmp/r2prof
mkdir -p $OUT $TMP
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --ntime 600 --no-config-d --no-e2e-api"
$CMD > $OUT/r2_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/r2_plain.log; exit 1; }
# 1. launch list of the graph-free config-A steps
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file $OUT/r2_launches_configA.csv $CMD > $TMP/launch.log 2>&1; echo "launch list rc=$?"
# 2. full capture of one config-A step (the kernels of the 6th step)
timeout 400 ncu --set full --clock-control none \
  -k regex:'conv3x3|dense|adam|sumsq|head_loss|build_batch|sgemm|skinny|outer' -s 215 -c 43 -o $TMP/configA -f \
  $CMD > $TMP/fullA.log 2>&1; echo "config A full capture rc=$?"
ncu -i $TMP/configA.ncu-rep --page raw --csv > $OUT/r2_prof_configA_raw.csv 2>/dev/null
# 3. config D (512x512, bf16, B=16): convolutions + dense GEMMs with the full set ...
DINCAE_NO_SIDE_STREAM=1 timeout 500 ncu --set full --clock-control none \
  -k regex:'conv3x3_bf16_kernel|conv3x3_wgrad_bf16_kernel|dense_tc_kernel' -s 68 -c 34 -o $TMP/configD -f \
  python tools/configD_step.py 16 bf16 > $TMP/fullD.log 2>&1; echo "config D conv capture rc=$?"
ncu -i $TMP/configD.ncu-rep --page raw --csv > $OUT/r2_prof_configD_raw.csv 2>/dev/null
# ... and its HBM-bound kernels with a handful of metrics (a full-set replay of the 8 GB Adam
# kernels takes minutes per kernel)
DINCAE_NO_SIDE_STREAM=1 timeout 400 ncu --clock-control none \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  -k regex:'adam_factored|gram_|sumsq_multi|adam_multi|build_batch|p4_to_p8|p8_to_p4|head_loss|conv_weights' -s 56 -c 28 --csv \
  --log-file $OUT/r2_prof_configD_hbm.csv python tools/configD_step.py 16 bf16 > $TMP/hbmD.log 2>&1; echo "config D hbm capture rc=$?"
ls -la $OUT/r2_prof_config* $OUT/r2_launches_configA.csv
tail -3 $TMP/fullD.log
