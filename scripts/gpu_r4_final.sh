#!/bin/bash
# round 4 evidence run on one B200: tests, smoke, bench (+ reference arm), ncu launch list, full captures of the top kernels
TAG=${1:-r4z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_$TAG.log; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log | cut -c1-200
timeout 400 python bench.py > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu-baseline --skip-roofline > gpurun_out/ncu_launch_$TAG.log 2>&1
for spec in "conv_wgrad_tc_kernel conv1_wgrad_mma" "conv_fwd_tc_kernel conv2_fwd_tc" "conv_fwd_tc_kernel conv1_fwd_tc"; do
  set -- $spec
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:$1 -s 2 -c 1 -f -o gpurun_out/prof_$2_$TAG \
    python scripts/bench_kernels.py --only $2 > gpurun_out/ncu_$2_$TAG.log 2>&1
done
ls gpurun_out | grep $TAG
