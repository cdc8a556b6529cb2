#!/bin/bash
# ncu evidence for the bench command (run only after the plain command exited 0).
# usage: tools/gpu_profile.sh <tag> <kernel-regex> [bench args...]
set -u
TAG=$1; KREGEX=$2; shift 2
ARGS="${*:---steps 3 --warmup 3 --no-cpu-baseline}"
mkdir -p gpurun_out
timeout 600 python bench.py $ARGS > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_$TAG.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 6 -c 6 \
    -o gpurun_out/prof_$TAG -f python bench.py $ARGS > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | tail -8
