#!/bin/bash
# round 5 on 8 GPUs: c3 weak scaling and c5 as configured (B=1024 over 8 GPUs) with the row-sweep kernels
TAG=${1:-r5n}
mkdir -p gpurun_out
for cfg in c3 c5; do
  NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --config $cfg --gpus 8 --steps 50 --warmup 5 --skip-cpu-baseline --skip-roofline > gpurun_out/bench_${TAG}_${cfg}_n8.log 2>&1
  tail -1 gpurun_out/bench_${TAG}_${cfg}_n8.log | cut -c1-300
done
