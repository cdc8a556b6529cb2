#!/bin/bash
# ncu --set full captures of named kernels of the bench command, one pass per kernel.
# usage: tools/gpu_full.sh <tag> <kernel-regex> [<kernel-regex> ...]
set -u
TAG=$1; shift
ARGS="--steps 2 --warmup 3 --no-cpu-baseline --no-extra"
export AREAD_GRAPHS=0      # eager launches: every kernel of a step is a launch ncu can select
mkdir -p gpurun_out
timeout 600 python bench.py $ARGS > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
for K in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s ${NCU_SKIP:-6} -c ${NCU_COUNT:-3} \
      -o gpurun_out/prof_${TAG}_$K -f python bench.py $ARGS > gpurun_out/ncu_full_${TAG}_$K.log 2>&1
  echo "$K exit $?"
done
