#!/bin/bash
# Runs on a multi-GPU box under `gpurun --gpus N`: the data-parallel bench at N (and 1) GPUs.
# usage: scripts/gpu_multi.sh <tag> <N>
TAG=${1:-m}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpus_$TAG.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpus_$TAG.txt 2>&1
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --skip-cpu-baseline > gpurun_out/bench_${TAG}_n1.log 2>&1; tail -1 gpurun_out/bench_${TAG}_n1.log | cut -c1-300
for n in 2 4 8; do
  if [ $n -le $N ]; then
    NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 50 --warmup 5 --skip-cpu-baseline > gpurun_out/bench_${TAG}_n$n.log 2>&1
    tail -1 gpurun_out/bench_${TAG}_n$n.log | cut -c1-300
  fi
done
