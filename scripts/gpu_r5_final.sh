#!/bin/bash
# round 5 evidence run on one B200: tests, smoke, bench, ncu launch list, full captures of the top kernels, graph timeline,
# compute-sanitizer over the new row-sweep kernels
TAG=${1:-r5z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_$TAG.log; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log | cut -c1-200
timeout 500 python bench.py > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu-baseline --skip-roofline > gpurun_out/ncu_launch_$TAG.log 2>&1
cap() {   # regex name bench-kernel
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$1" -s 2 -c 1 -f \
    -o gpurun_out/prof_$2_$TAG python scripts/bench_kernels.py --only $3 > gpurun_out/ncu_$2_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_$2_$TAG.log | cut -c1-120
}
cap 'conv_row_tc_kernel<\(int\)5>' conv2_fwd_row conv2_fwd_tc
cap 'conv_row_tc_kernel<\(int\)5>' conv2_dgrad_row conv2_dgrad_tc
cap 'conv_wgrad_row_kernel<\(int\)5>' conv2_wgrad_row conv2_wgrad_mma
cap 'conv_fwd_tc_kernel<\(int\)5, \(int\)1>' conv1_fwd_tc conv1_fwd_tc
cap 'conv_wgrad_tc_kernel' conv1_wgrad_tc conv1_wgrad_mma
TRACE_MODE=graph timeout 120 python scripts/trace_step.py 4 > gpurun_out/trace_$TAG.txt 2>&1
SEL='(test_conv_tc_from_pieces and (8-32-32-5-1 or 7-16-16-3-1 or 9-6-4-5-1 or 8-16-16-3-3)) or (test_conv_dgrad_tc and (8-16-16-3-1 or 5-32-48-5-1 or 8-32-32-5-5)) or (test_conv_wgrad_row and (8x32x32_k5 or 7x16x16_k3 or 9x6x4_k5))'
for TOOL in memcheck racecheck; do
  timeout 300 compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_kernels.py -x -q -k "$SEL" > gpurun_out/sanitizer_${TOOL}_$TAG.log 2>&1
  echo "$TOOL exit $?" | tee -a gpurun_out/sanitizer_${TOOL}_$TAG.log
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_${TOOL}_$TAG.log | tail -3
done
ls gpurun_out | grep $TAG | tr '\n' ' '
