#!/bin/bash
# Multi-GPU bench on one box, the way the driver launches it.  Usage (under gpurun --gpus N): tools/scale_round.sh TAG N [extra bench args]
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 "$@" \
  > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err
echo "bench $N gpus rc=$?"; cut -c1-300 gpurun_out/bench_${N}gpu_$TAG.json; tail -3 gpurun_out/bench_${N}gpu_$TAG.err
