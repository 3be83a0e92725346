#!/bin/bash
# round 5: full GPU suite, short bench, the other BASELINE configs on one GPU
TAG=${1:-r5l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/pytest_gpu_$TAG.log
timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu-baseline --skip-roofline > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log | cut -c1-230
timeout 300 python scripts/bench_configs.py 2>&1 | grep -v input_layer > gpurun_out/configs_$TAG.jsonl; cat gpurun_out/configs_$TAG.jsonl
