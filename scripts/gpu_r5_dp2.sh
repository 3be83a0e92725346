#!/bin/bash
# round 5 on 2 GPUs: replica parity test (P2P / NCCL / host transports) and the data-parallel bench lines with the row-sweep kernels
TAG=${1:-r5m}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpus_$TAG.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dp.py -x -q -s > gpurun_out/pytest_dp_$TAG.log 2>&1; tail -2 gpurun_out/pytest_dp_$TAG.log
for cfg in c3 c4; do
  NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --config $cfg --gpus 2 --steps 50 --warmup 5 --skip-cpu-baseline --skip-roofline > gpurun_out/bench_${TAG}_${cfg}_n2.log 2>&1
  tail -1 gpurun_out/bench_${TAG}_${cfg}_n2.log | cut -c1-300
done
