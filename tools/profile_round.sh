#!/bin/bash
# One gpurun call that produces everything profiles/ needs for a round (run from the repo root):
#   gpurun --timeout 900 -- 'bash tools/profile_round.sh rN'
# 1. GPU test suite, 2. bench line (with CPU baseline), 3. ncu launch list of a 3-step run,
# 4. ncu --set full of the convolution kernels of one step (after 3 exited 0).
# Outputs land in gpurun_out/; copy what should be judged into profiles/ (tools/step_launches.py,
# tools/ncu_pick.py turn the csv / report into the tables of profiles/*_summary_*.md).
set -u
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --ntime 600"
$CMD > $OUT/plain_${TAG}.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_launch_${TAG}.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:'conv3x3_wgrad_tc_kernel|conv3x3_tc_kernel' -s 30 -c 23 -o $OUT/prof_tc_${TAG} -f \
  $CMD > $OUT/ncu_full_${TAG}.log 2>&1; echo "full capture rc=$?"
DINCAE_NO_SIDE_STREAM=1 timeout 90 python tools/configD_step.py 8 2>&1 | tail -45 > $OUT/configD_${TAG}.txt
