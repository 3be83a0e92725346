#!/bin/bash
# Runs on the GPU box under gpurun: tests, smoke, bench, ncu launch list + full capture of the top kernels.
# usage: scripts/gpu_round.sh [tag] [kernel-regexes for the full capture...]
TAG=${1:-r1}; shift
KERNELS=${@:-conv_fwd_tc_kernel conv_wgrad_kernel}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | tail -400 > gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log
# every launch of 2 steady-state steps with device time (serialised, cold cache: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu-baseline --skip-roofline > gpurun_out/ncu_launch_$TAG.log 2>&1
# full capture of the top kernels (one launch each)
for K in $KERNELS; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o gpurun_out/prof_${K}_$TAG \
    python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu-baseline --skip-roofline > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
ls -la gpurun_out | tail -20
