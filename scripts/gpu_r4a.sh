#!/bin/bash
# round 4: tcgen05 weight gradient - kernel timing, whole-step parity, bench, ncu launch list
TAG=${1:-r4a}
mkdir -p gpurun_out
timeout 200 python scripts/bench_kernels.py --only conv1_wgrad_mma > gpurun_out/kernels_$TAG.jsonl 2>&1; tail -2 gpurun_out/kernels_$TAG.jsonl | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_step_pinned.py tests/test_gpu_nets.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_$TAG.log; cat gpurun_out/pytest_$TAG.log
timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu-baseline > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log | cut -c1-600
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu-baseline --skip-roofline > gpurun_out/ncu_launch_$TAG.log 2>&1
grep -c . gpurun_out/launches_$TAG.csv
