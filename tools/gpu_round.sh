#!/bin/bash
# One GPU-box visit: parity tests, smoke, the bench line, then the ncu launch list of a short bench run.
# usage: tools/gpu_round.sh <tag>        (PYTEST_ARGS / BENCH_ARGS / SKIP_TESTS / SKIP_NCU from the environment)
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
if [ -z "${SKIP_TESTS:-}" ]; then
  eval "timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ${PYTEST_ARGS:-}" > gpurun_out/pytest_$TAG.log 2>&1
  echo "pytest exit $?" | tee -a gpurun_out/pytest_$TAG.log
  tail -30 gpurun_out/pytest_$TAG.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1
  echo "smoke exit $?"; tail -3 gpurun_out/smoke_$TAG.log
fi
timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; tail -1 gpurun_out/bench_$TAG.log | cut -c1-1500; tail -5 gpurun_out/bench_$TAG.err
if [ -z "${SKIP_NCU:-}" ]; then
  PARGS="--steps 2 --warmup 3 --no-cpu-baseline --no-extra"
  timeout 300 python bench.py $PARGS > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
  # eager launches (no graph replay) so that every kernel of a step is listed; steps 4-5 are the two timed ones
  AREAD_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
      --log-file gpurun_out/launches_$TAG.csv python bench.py $PARGS > gpurun_out/ncu_launch_$TAG.log 2>&1
  echo "launch list exit $?"
  python tools/summarize_ncu.py launches gpurun_out/launches_$TAG.csv 4 5 > gpurun_out/launches_$TAG.txt 2>&1
  head -45 gpurun_out/launches_$TAG.txt
fi
