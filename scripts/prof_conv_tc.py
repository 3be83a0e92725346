#!/usr/bin/env python
"""Where does a conv_fwd_tc CTA spend its cycles?  Needs the diagnosis build:
  CARTPOLEPP_NVCC_EXTRA=-DCONV_TC_PROF python -c "import __graft_entry__ as g; g.build(force=True)"
then  python scripts/prof_conv_tc.py [conv1_fwd_tc|conv2_fwd_tc|conv3_fwd_tc|conv2_dgrad_tc|conv3_dgrad_tc]
prints per-role wait cycles (mean / max over the CTAs of the LAST launch).  Numbers under this build are not bench values."""
import ctypes as C
import subprocess
import sys
import os
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
NAMES = ["prologue", "mma total", "mma waits FULL_PL (fill-bound)", "mma waits EMPTY_ACC (epilogue-bound)", "fill total",
         "fill waits EMPTY_PL (mma-bound)", "fill waits STAGE (TMA latency)", "epi total", "epi waits FULL_ACC (mma-bound)",
         "kernel", "tiles", "units"]


def main():
  which = sys.argv[1:] or ["conv1_fwd_tc"]
  from cartpoleplusplus_b200 import _lib as L
  import torch
  lib = L.lib()
  fn = lib.cpp_debug_conv_tc_prof
  skip = int(os.environ.get("CONV_TC_SKIP", "0"))
  if skip:
    assert lib.cpp_debug_conv_tc_skip(skip) == 0
    print("#### skip mask %d (1 no MMAs, 2 no re-layout, 4 no epilogue math): timing experiment, results are wrong" % skip)
  for name in which:
    # run the kernel through bench_kernels' set-up in this process so that g_prof holds its last launch
    import scripts.bench_kernels as bk
    sys.argv = ["bench_kernels.py", "--only", name, "--reps", "3"]
    bk.main()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * (160 * 12))()
    assert fn(buf) == 0
    a = np.frombuffer(buf, dtype=np.uint64).reshape(160, 12).astype(np.float64)
    a = a[a[:, 9] > 0]
    print("== %s: %d CTAs" % (name, a.shape[0]))
    for i, n in enumerate(NAMES):
      print("  %-42s mean %9.0f   min %9.0f   max %9.0f" % (n, a[:, i].mean(), a[:, i].min(), a[:, i].max()))


if __name__ == "__main__":
  main()
