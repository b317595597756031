#!/bin/bash
# Usage (under gpurun): bash profiles/run_ncu.sh <tag> <kernel-regex> [skip] [count]
set -u
TAG=$1; KR=$2; SKIP=${3:-4}; CNT=${4:-2}
OUT=gpurun_out
SHORT="python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --e2e-steps 1"
$SHORT > $OUT/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_$TAG.csv $SHORT > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch-list rc=$?"
$SHORT > $OUT/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$KR -s $SKIP -c $CNT -o $OUT/prof_$TAG -f $SHORT > $OUT/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
