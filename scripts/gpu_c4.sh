#!/bin/bash
# BASELINE config c4 (NAF, 64x64x18, batch 512) as configured on 4 GPUs, and on 1 GPU for the scaling ratio
TAG=${1:-m}
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --config c4 --gpus 4 --steps 50 --warmup 5 --skip-cpu-baseline --skip-roofline > gpurun_out/bench_${TAG}_c4_n4.log 2>&1
tail -1 gpurun_out/bench_${TAG}_c4_n4.log | cut -c1-330
timeout 600 python bench.py --config c4 --gpus 1 --steps 50 --warmup 5 --skip-cpu-baseline --skip-roofline > gpurun_out/bench_${TAG}_c4_n1.log 2>&1
tail -1 gpurun_out/bench_${TAG}_c4_n1.log | cut -c1-330
