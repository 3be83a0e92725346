#!/usr/bin/env python
"""Where does a conv_wgrad_tc CTA spend its cycles?  Needs the diagnosis build:
  CARTPOLEPP_NVCC_EXTRA=-DWGTC_PROF python -c "import __graft_entry__ as g; g.build(force=True)"
prints per-role wait cycles (mean / min / max over the CTAs of the LAST launch).  Numbers under this build are not bench values."""
import ctypes as C
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
NAMES = ["kernel", "mma loop", "mma waits FULL (fill-bound)", "mma waits EMPTY_ACC (epilogue-bound)", "warp 0 loop",
         "warp 0 (X items) waits FREE (mma-bound)", "warp 0 waits STAGE_FULL (TMA)", "warp 0 item work", "warp 4 (dY items) waits FREE",
         "producer waits STAGE_FREE", "warp 4 item work", "warp 4 waits STAGE_FULL"]


def main():
  from cartpoleplusplus_b200 import _lib as L
  import torch
  lib = L.lib()
  import scripts.bench_kernels as bk
  if len(sys.argv) > 2:
    bk.B = int(sys.argv[2])          # batch (default 256): kernel cycles vs number of steps separates fixed from per-step cost
  which = sys.argv[1] if len(sys.argv) > 1 else "conv1_wgrad_mma"
  sys.argv = ["bench_kernels.py", "--only", which, "--reps", "3"]
  bk.main()
  torch.cuda.synchronize()
  buf = (C.c_ulonglong * (160 * 12))()
  assert lib.cpp_debug_wgrad_tc_prof(buf) == 0
  a = np.frombuffer(buf, dtype=np.uint64).reshape(160, 12).astype(np.float64)
  a = a[a[:, 0] > 0]
  print("== conv_wgrad_tc (%s): %d CTAs" % (which, a.shape[0]))
  for i, n in enumerate(NAMES):
    print("  %-42s mean %9.0f   min %9.0f   max %9.0f" % (n, a[:, i].mean(), a[:, i].min(), a[:, i].max()))


if __name__ == "__main__":
  main()
