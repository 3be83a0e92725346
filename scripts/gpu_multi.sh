#!/bin/bash
# Runs on a multi-GPU box under `gpurun --gpus N`: the 2-rank NCCL parity test, then the data-parallel bench.
# usage: scripts/gpu_multi.sh <tag> <N> [configs...]     (configs default: c3)
TAG=${1:-m}; N=${2:-2}; shift; shift
CONFIGS=${@:-c3}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpus_$TAG.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpus_$TAG.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dp.py -x -q -s > gpurun_out/pytest_dp_$TAG.log 2>&1; tail -3 gpurun_out/pytest_dp_$TAG.log
run() {   # config n extra-flags...
  local cfg=$1 n=$2; shift; shift
  local out=gpurun_out/bench_${TAG}_${cfg}_n$n$(echo "$@" | tr -d ' -').log
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --config $cfg --gpus 1 --steps 50 --warmup 5 --skip-cpu-baseline --skip-roofline "$@" > $out 2>&1
  else
    NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --config $cfg --gpus $n --steps 50 --warmup 5 --skip-cpu-baseline --skip-roofline "$@" > $out 2>&1
  fi
  tail -1 $out | cut -c1-400
}
for cfg in $CONFIGS; do
  for n in 1 2 4 8; do
    if [ $n -le $N ]; then
      run $cfg $n
      if [ $n -gt 1 ] && [ "$cfg" = "c3" ]; then run $cfg $n --host-allreduce; fi
      if [ "$cfg" = "c3" ] && [ $n -gt 1 ]; then run $cfg $n --scaling strong; fi
    fi
  done
done
