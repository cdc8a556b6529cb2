#!/bin/bash
# One GPU-box visit: parity tests, smoke, a short bench, and (optionally) the ncu launch list.
# Everything is logged under gpurun_out/ ; every step has its own timeout.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py ${BENCH_ARGS:---steps 10 --warmup 3} > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -2 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
