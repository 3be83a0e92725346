#!/bin/bash
# Runs on the GPU box under gpurun: compute-sanitizer memcheck / racecheck / synccheck over the small-shape kernel tests
# (the warp-specialised mbarrier / TMEM kernels of conv_tc.cu, the mma.sync weight gradient, the FC kernels).
# usage: scripts/gpu_sanitize.sh <tag> [per-tool timeout seconds]
TAG=${1:-s}; TMO=${2:-420}
mkdir -p gpurun_out
SEL='(test_conv_tc_forward and (8-64-64-9 or 3-33-31 or 4-32-32-10)) or (test_conv_tc_from_pieces and (8-32-32 or 3-12-12)) or (test_conv_dgrad_tc and (8-16-16 or 3-25-25 or 1-4-4)) or (test_conv_wgrad_mma and (8x64x64x9_k5_n2 or 3x33x31 or 8x16x16x10_k3_n1_p2 or 5x8x8)) or (test_conv_layer and (5x8x8 or 7x12x12))'
for TOOL in memcheck racecheck synccheck; do
  timeout $TMO compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_kernels.py -x -q -k "$SEL" \
    > gpurun_out/sanitizer_${TOOL}_$TAG.log 2>&1
  echo "$TOOL exit $?" | tee -a gpurun_out/sanitizer_${TOOL}_$TAG.log
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_${TOOL}_$TAG.log | tail -3
done
# whole-step (streams + graph) under memcheck at a small size
timeout $TMO compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_nets.py -x -q -k "fused_step_streams or (golden and ddpg_pixel-tensorcore) or (naf_golden and naf_pixel-tensorcore)" \
  > gpurun_out/sanitizer_step_$TAG.log 2>&1
echo "step memcheck exit $?" | tee -a gpurun_out/sanitizer_step_$TAG.log
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_step_$TAG.log | tail -3
