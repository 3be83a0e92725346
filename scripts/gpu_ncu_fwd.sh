#!/bin/bash
# full ncu captures of the conv1 (<5,1>) and conv2 (<5,0>) instantiations of conv_fwd_tc_kernel
TAG=${1:-x}
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_fwd_tc_kernel<\(int\)5, \(int\)0>' -s 2 -c 1 -f \
  -o gpurun_out/prof_conv2_fwd_tc_$TAG python scripts/bench_kernels.py --only conv2_fwd_tc > gpurun_out/ncu_conv2_fwd_tc_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_fwd_tc_kernel<\(int\)5, \(int\)1>' -s 2 -c 1 -f \
  -o gpurun_out/prof_conv1_fwd_tc_$TAG python scripts/bench_kernels.py --only conv1_fwd_tc > gpurun_out/ncu_conv1_fwd_tc_$TAG.log 2>&1
tail -2 gpurun_out/ncu_conv1_fwd_tc_$TAG.log | cut -c1-120
