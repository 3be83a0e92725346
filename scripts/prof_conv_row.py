#!/usr/bin/env python
"""Where does a conv_row_tc CTA spend its cycles?  Needs the diagnosis build:
  CARTPOLEPP_NVCC_EXTRA=-DCONV_ROW_PROF python -c "import __graft_entry__ as g; g.build(force=True)"
then  CONV_ROW=1|3 python scripts/prof_conv_row.py [conv2_fwd_tc|conv3_fwd_tc|conv2_dgrad_tc|conv3_dgrad_tc]
prints per-role cycles (mean / min / max over the CTAs of the LAST launch).  Numbers under this build are not bench values."""
import ctypes as C
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
NAMES = ["prologue", "mma total", "mma waits FULL_A (producer-bound)", "mma waits ACC_FREE (epilogue-bound)", "producer total",
         "producer waits EMPTY_A (mma-bound)", "producer waits copies", "epi total", "epi waits ACC_FULL (mma-bound)", "kernel", "rows", "tiles"]


WNAMES = ["prologue", "mma total", "mma waits DFULL (dY producers)", "mma waits XFULL (TMA)", "mma waits ACC_EMPTY (epilogue)",
          "dY group 0 total", "dY group 0 waits DONE (mma-bound)", "epi total", "epi waits ACC_FULL", "kernel", "steps", "X steps"]


def main():
  which = sys.argv[1:] or ["conv2_fwd_tc"]
  from cartpoleplusplus_b200 import _lib as L
  import torch
  lib = L.lib()
  for name in which:
    wg = "wgrad" in name                       # -DWGROW_PROF build, WGRAD_TC=5
    fn = lib.cpp_debug_wgrad_row_prof if wg else lib.cpp_debug_conv_row_prof
    import scripts.bench_kernels as bk
    sys.argv = ["bench_kernels.py", "--only", name, "--reps", "3"]
    bk.main()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * (160 * 12))()
    assert fn(buf) == 0
    a = np.frombuffer(buf, dtype=np.uint64).reshape(160, 12).astype(np.float64)
    a = a[a[:, 9] > 0]
    print("== %s (CONV_ROW=%s): %d CTAs" % (name, os.environ.get("CONV_ROW", "default"), a.shape[0]))
    for i, n in enumerate(WNAMES if wg else NAMES):
      print("  %-42s mean %9.0f   min %9.0f   max %9.0f" % (n, a[:, i].mean(), a[:, i].min(), a[:, i].max()))


if __name__ == "__main__":
  main()
