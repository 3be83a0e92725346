#!/usr/bin/env python
"""A/B timing of a cpp_set_option switch on ONE box, interleaved so that clock / box differences cancel:
  python scripts/ab_option.py prep_hoist [c3|c5] [rounds] [pinned|computed] [values ...]
prints the median ms/step of the fused DDPG step (device-resident batch, CUDA events) per option value (default 0 and 1);
`pinned`: whitening statistics handed in, as bench.py does."""
import json
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import gpu_util as U                     # noqa: E402
from oracle.make_golden import _batch               # noqa: E402
from cartpoleplusplus_b200 import _lib as L         # noqa: E402


def main():
  name = sys.argv[1]
  cfg = sys.argv[2] if len(sys.argv) > 2 else "c3"
  rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 7
  shape, B = {"c3": ((64, 64, 3, 1, 3), 256), "c5": ((128, 128, 3, 2, 4), 128)}[cfg]
  rs = np.random.RandomState(0)
  nets, eng, o = U.make_ddpg(shape, True, None, batch_size=B)
  b = U.Batch(*[torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in _batch(rs, B, shape)])
  moments = None
  if len(sys.argv) > 4 and sys.argv[4] == "pinned":        # whitening statistics handed in, as bench.py does (per-slot sums)
    import ctypes as C
    lib = L.lib()
    cin = int(np.prod(shape[2:]))
    ms = []
    for x in (b.state_1, b.state_2):
      scratch = torch.zeros(int(lib.cpp_moments_scratch_doubles(cin)), dtype=torch.float64, device="cuda")
      mi = torch.zeros(2 * cin, dtype=torch.float32, device="cuda")
      L.check(lib.cpp_channel_moments(L.ptr(x), 1, C.c_int64(B * shape[0] * shape[1]), cin, L.ptr(scratch), L.ptr(mi), L.stream_ptr()))
      ms.append(mi)
    moments = tuple(ms)
  step = (lambda: eng.train_step(b, moments=moments)) if moments is not None else (lambda: eng.train_step(b))
  values = [int(v) for v in sys.argv[5:]] or [0, 1]
  res = {v: [] for v in values}
  for r in range(rounds):
    for v in values:
      L.check(L.lib().cpp_set_option(name.encode(), v))
      for _ in range(5):
        step()
      torch.cuda.synchronize()
      a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      for _ in range(40):
        step()
      e.record(); torch.cuda.synchronize()
      res[v].append(a.elapsed_time(e) / 40)
  print(json.dumps({"option": name, "config": cfg, "median_ms_per_step": {str(v): round(float(np.median(res[v])), 4) for v in values},
                    "all": {str(v): [round(x, 4) for x in res[v]] for v in values}}))


if __name__ == "__main__":
  main()
