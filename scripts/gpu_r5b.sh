#!/bin/bash
# round 5: full GPU test suite + bench line + graph timeline
TAG=${1:-r5f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/pytest_gpu_$TAG.log
timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu-baseline --skip-roofline > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log | cut -c1-260
TRACE_MODE=graph timeout 120 python scripts/trace_step.py 4 > gpurun_out/trace_$TAG.txt 2>&1; grep -v input_layer gpurun_out/trace_$TAG.txt | tail -38
