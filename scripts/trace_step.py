#!/usr/bin/env python
"""Prints the fork/join timeline of fused DDPG steps at c3 (eager multi-stream mode, CARTPOLEPP_TRACE=1, graphs off)."""
import os, sys
mode = os.environ.get("TRACE_MODE", "eager")       # eager | graph
os.environ["CARTPOLEPP_TRACE"] = "2" if mode == "graph" else "1"; os.environ["CARTPOLEPP_GRAPHS"] = "1" if mode == "graph" else "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import gpu_util as U
from oracle.make_golden import ddpg_params, _batch
shape, B = (64, 64, 3, 1, 3), 256
rs = np.random.RandomState(1)
P = ddpg_params(rs, shape, True)
batch = U.Batch(*[torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in _batch(rs, B, shape)])
nets, eng, o = U.make_ddpg(shape, True, {k: v.numpy() for k, v in P.items()}, batch_size=B)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
  eng.train_step(batch)
torch.cuda.synchronize()
