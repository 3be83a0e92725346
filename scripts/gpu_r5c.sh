#!/bin/bash
TAG=${1:-r5h}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "wgrad_row" 2>&1 | tail -25 > gpurun_out/pytest_wrow_$TAG.log
cat gpurun_out/pytest_wrow_$TAG.log
for w in 5 1; do
  WGRAD_TC=$w timeout 100 python scripts/bench_kernels.py --only conv2_wgrad_mma,conv3_wgrad_mma 2>&1 | grep -v input_layer | cut -c1-60,150-260
done
