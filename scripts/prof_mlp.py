#!/usr/bin/env python
"""Phase timestamps of mlp_forward_kernel (needs the -DMLP_PROF build): cycles from kernel start to: input staged, start of each
layer, end - mean over the CTAs of the LAST launch.  usage: python scripts/prof_mlp.py [fc_actor_fwd|fc_critic_fwd]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
def main():
  which = sys.argv[1:] or ["fc_actor_fwd"]
  from cartpoleplusplus_b200 import _lib as L
  import torch
  lib = L.lib()
  for name in which:
    import scripts.bench_kernels as bk
    sys.argv = ["bench_kernels.py", "--only", name, "--reps", "3"]
    bk.main(); torch.cuda.synchronize()
    buf = (C.c_longlong * (64 * 8))()
    assert lib.cpp_debug_mlp_prof(buf) == 0
    a = np.frombuffer(buf, dtype=np.int64).reshape(64, 8).astype(np.float64)
    a = a[a[:, 0] > 0]
    d = a - a[:, :1]
    print("== %s: %d CTAs; cycles since kernel start (mean): x staged %.0f | layer starts %s | end %.0f" % (
        name, a.shape[0], d[:, 1].mean(), [int(d[:, i].mean()) for i in range(2, 6)], d[:, 7].mean()))
if __name__ == "__main__":
  main()
