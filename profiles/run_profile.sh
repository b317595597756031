#!/bin/bash
# Usage (under gpurun): bash profiles/run_profile.sh <tag>
# 1. full bench line  2. launch list (ncu, cold-cache serialised: compare SHARES)  3. ncu --set full of the RoIAlign kernels
set -u
TAG=${1:-r1}
OUT=gpurun_out
SHORT="python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-e2e"
python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err; echo "ref rc=$?"
$SHORT > $OUT/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $SHORT > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch-list rc=$?"
$SHORT > $OUT/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:roi_align -s 6 -c 2 -o $OUT/prof_roialign_$TAG -f $SHORT > $OUT/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
tail -c 600 $OUT/bench_$TAG.err
