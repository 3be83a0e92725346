#!/bin/bash
# quick GPU iteration: agent-level parity tests + bench line (+ optional other configs)
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_conv_tc.py tests/test_saver.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_quick_$TAG.log
timeout 200 python bench.py --steps 50 --warmup 5 --skip-cpu-baseline > gpurun_out/bench_$TAG.log 2>&1
timeout 200 python scripts/bench_configs.py > gpurun_out/configs_$TAG.jsonl 2>&1
cat gpurun_out/pytest_quick_$TAG.log; tail -1 gpurun_out/bench_$TAG.log | cut -c1-300; grep -v input_layer gpurun_out/configs_$TAG.jsonl | cut -c1-200
