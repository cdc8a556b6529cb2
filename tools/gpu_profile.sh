#!/bin/bash
# ncu evidence for the bench command (run only after the plain command exited 0).
# usage: tools/gpu_profile.sh <tag> <kernel-regex> <skip> <count> [bench args...]
set -u
TAG=$1; KREGEX=$2; SKIP=$3; COUNT=$4; shift 4
ARGS="${*:---steps 2 --warmup 3 --no-cpu-baseline}"
mkdir -p gpurun_out
timeout 600 python bench.py $ARGS > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_$TAG.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $COUNT --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 4 -c 4 \
    -o gpurun_out/prof_$TAG -f python bench.py $ARGS > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
