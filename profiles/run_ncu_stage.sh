#!/bin/bash
# Usage (under gpurun): bash profiles/run_ncu_stage.sh <tag> <kernel-regex> [batch]
set -u
TAG=$1; KR=$2; B=${3:-2}
OUT=gpurun_out
CMD="python profiles/stage_bench.py $B 5"
$CMD > $OUT/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch-list rc=$?"
$CMD > $OUT/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$KR -s 12 -c 6 -o $OUT/prof_$TAG -f $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
cat $OUT/plain_$TAG.log
