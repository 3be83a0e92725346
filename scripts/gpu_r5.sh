#!/bin/bash
# round 5 iteration: row-sweep conv kernel parity + kernel timings (both routes)
TAG=${1:-r5a}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q -x -k "from_pieces or dgrad_tc" 2>&1 | tail -25 > gpurun_out/pytest_row_$TAG.log
cat gpurun_out/pytest_row_$TAG.log
for r in 1 3 0; do
  CONV_ROW=$r timeout 200 python scripts/bench_kernels.py --only conv2_fwd_tc,conv2_dgrad_tc,conv3_fwd_tc,conv3_dgrad_tc 2>&1 | grep -v input_layer | cut -c1-400 > gpurun_out/kernels_row${r}_$TAG.jsonl
  python - <<PY
import json
for l in open("gpurun_out/kernels_row${r}_$TAG.jsonl"):
  try:
    d = json.loads(l); print("row=$r", d["kernel"], round(d["us_median"], 1), "us")
  except Exception: print(l.strip()[:300])
PY
done
