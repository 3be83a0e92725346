#!/bin/bash
# conv_tc iteration loop on the GPU box: role-wait diagnosis (PROF build), then the normal build: parity tests + kernel timings (+ bench)
TAG=${1:-x}
mkdir -p gpurun_out
CARTPOLEPP_NVCC_EXTRA=-DCONV_TC_PROF python -c "import __graft_entry__ as g; g.build(force=True)" > /dev/null 2>&1
timeout 200 python scripts/prof_conv_tc.py conv1_fwd_tc 2>&1 | grep -v input_layer > gpurun_out/prof_roles_$TAG.txt
python -c "import __graft_entry__ as g; g.build(force=True)" > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_nets.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_convtc_$TAG.log
timeout 200 python scripts/bench_kernels.py > gpurun_out/kernels_$TAG.jsonl 2>&1
timeout 200 python bench.py --steps 50 --warmup 5 --skip-cpu-baseline > gpurun_out/bench_$TAG.log 2>&1
timeout 200 python scripts/bench_configs.py > gpurun_out/configs_$TAG.jsonl 2>&1
cat gpurun_out/prof_roles_$TAG.txt; cat gpurun_out/pytest_convtc_$TAG.log; cut -c1-150 gpurun_out/kernels_$TAG.jsonl; tail -1 gpurun_out/bench_$TAG.log | cut -c1-300; grep -v input_layer gpurun_out/configs_$TAG.jsonl | cut -c1-200
