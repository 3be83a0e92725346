#!/usr/bin/env python
"""Step times of the other BASELINE.json configs on ONE B200 (per-GPU shard of the data-parallel configs), device-resident
synthetic batches, CUDA events, 20 steps after 5 warm-up steps.  Not the bench.py headline (that is config 3) - a record of
how the same code paths behave at the other shapes.

  c1 LRPG low-dim (2000 observations)   c2 DDPG low-dim B=256   c3 DDPG 64x64x9 B=256
  c4 NAF 64x64x18, shard B=128 (512 over 4 GPUs)                c5 DDPG 128x128x24, shard B=128 (1024 over 8 GPUs)"""
import json
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import gpu_util as U                     # noqa: E402
from oracle.make_golden import _batch               # noqa: E402


def timed(fn, steps=20, warmup=5):
  for _ in range(warmup):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(steps):
    fn()
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) / steps


def dev_batch(rs, B, shape):
  return U.Batch(*[torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in _batch(rs, B, shape)])


def main():
  rs = np.random.RandomState(0)
  out = []
  # c2
  nets, eng, o = U.make_ddpg((2, 2, 7), False, None, batch_size=256)
  b = dev_batch(rs, 256, (2, 2, 7))
  out.append(dict(config="c2 DDPG low-dim B=256", ms_per_step=timed(lambda: eng.train_step(b))))
  # c3
  nets, eng, o = U.make_ddpg((64, 64, 3, 1, 3), True, None, batch_size=256)
  b = dev_batch(rs, 256, (64, 64, 3, 1, 3))
  out.append(dict(config="c3 DDPG 64x64x9 B=256", ms_per_step=timed(lambda: eng.train_step(b))))
  # c4 shard
  naf, nets, eng, o = U.make_naf((64, 64, 3, 2, 3), True, None, batch_size=128, optimiser="Adam", optimiser_args={"learning_rate": 1e-4})
  b4 = dev_batch(rs, 128, (64, 64, 3, 2, 3))
  out.append(dict(config="c4 NAF 64x64x18 shard B=128", ms_per_step=timed(lambda: eng.train(b4))))
  # c5 shard
  nets, eng, o = U.make_ddpg((128, 128, 3, 2, 4), True, None, batch_size=128)
  b5 = dev_batch(rs, 128, (128, 128, 3, 2, 4))
  out.append(dict(config="c5 DDPG 128x128x24 shard B=128", ms_per_step=timed(lambda: eng.train_step(b5))))
  for r in out:
    r["steps_per_s"] = 1e3 / r["ms_per_step"]
    print(json.dumps(r))


if __name__ == "__main__":
  main()
