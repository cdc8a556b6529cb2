#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): correctness of the row-sharded table (tools/check_sharded.py under torchrun),
# then the bench at N GPUs (and at fewer, when asked).  usage: tools/gpu_multi.sh <tag> <N> [bench worlds...]
set -u
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpu_$TAG.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $RUN --nproc-per-node $N --master-port 29511 tools/check_sharded.py > gpurun_out/sharded_${TAG}_w$N.log 2>&1
echo "check_sharded world=$N exit $?"; tail -3 gpurun_out/sharded_${TAG}_w$N.log
if [ -n "${DENSE_TOO:-}" ]; then
  AREAD_DENSE_GRAD_EXCHANGE=1 timeout 600 $RUN --nproc-per-node $N --master-port 29512 tools/check_sharded.py > gpurun_out/sharded_${TAG}_w${N}_dense.log 2>&1
  echo "check_sharded (dense reduce-scatter) world=$N exit $?"
fi
for W in "$@"; do
  if [ "$W" = "1" ]; then
    timeout 900 python bench.py --gpus 1 ${BENCH_ARGS:---no-cpu-baseline --no-extra} > gpurun_out/bench_${TAG}_w1.log 2> gpurun_out/bench_${TAG}_w1.err
  else
    timeout 900 $RUN --nproc-per-node $W --master-port 2952$W bench.py --gpus $W ${BENCH_ARGS:---no-cpu-baseline --no-extra} > gpurun_out/bench_${TAG}_w$W.log 2> gpurun_out/bench_${TAG}_w$W.err
    if [ -n "${DENSE_TOO:-}" ]; then
      AREAD_DENSE_GRAD_EXCHANGE=1 timeout 900 $RUN --nproc-per-node $W --master-port 2953$W bench.py --gpus $W ${BENCH_ARGS:---no-cpu-baseline --no-extra} > gpurun_out/bench_${TAG}_w${W}_dense.log 2> gpurun_out/bench_${TAG}_w${W}_dense.err
    fi
  fi
  echo "bench world=$W exit $?"; tail -1 gpurun_out/bench_${TAG}_w$W.log | cut -c1-400; tail -3 gpurun_out/bench_${TAG}_w$W.err
done
