#!/bin/bash
# Runs on an 8-GPU box under `gpurun --gpus 8`: the BASELINE configs as configured (c4 NAF B=512 on 4 GPUs, c5 DDPG 128x128
# B=1024 on 8 GPUs), c3 weak scaling at 8, and the 2-rank NCCL/P2P replica parity test.
TAG=${1:-m}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpus_$TAG.txt 2>&1
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests/test_gpu_dp.py -x -q -s > gpurun_out/pytest_dp_$TAG.log 2>&1; tail -2 gpurun_out/pytest_dp_$TAG.log
run() {   # config n extra-flags...
  local cfg=$1 n=$2; shift; shift
  local out=gpurun_out/bench_${TAG}_${cfg}_n$n$(echo "$@" | tr -d ' -').log
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --config $cfg --gpus 1 --steps 50 --warmup 5 --skip-cpu-baseline --skip-roofline "$@" > $out 2>&1
  else
    NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --config $cfg --gpus $n --steps 50 --warmup 5 --skip-cpu-baseline --skip-roofline "$@" > $out 2>&1
  fi
  tail -1 $out | cut -c1-330
}
run c5 8
run c4 4
run c3 8
run c3 1
run c5 1
run c4 1
