#!/usr/bin/env python
"""The gradient exchange alone (csrc/comm.cu) at the c3 buffer size, per transport: correctness of the sum against
torch.distributed, bit-identical results on every rank, and us per call back to back (CUDA events, max over ranks).
  torchrun --nproc-per-node N scripts/bench_allreduce.py"""
import json
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cartpoleplusplus_b200 import _lib as L, dp as dpmod      # noqa: E402
from tests import gpu_util as U                               # noqa: E402

dp = dpmod.DataParallel()
torch.cuda.set_device(dp.local_rank)
lib = L.lib()
out = {}
for transport in ("p2p", "nccl"):
  nets, eng, o = U.make_ddpg((64, 64, 3, 1, 3), True, None, batch_size=8)
  eng.set_data_parallel(dp, transport=transport)
  g = eng.buffers["grads"]
  gen = torch.Generator(device="cuda"); gen.manual_seed(100 + dp.rank)
  errs = []
  for it in range(5):
    g.copy_(torch.randn(g.numel(), device="cuda", generator=gen))
    want = g.double().clone()
    dist.all_reduce(want)
    L.check(lib.cpp_ddpg_all_reduce_grads(eng.handle, L.stream_ptr()))
    errs.append(float((g.double() - want).abs().max() / want.abs().max()))
    ref = g.clone(); dist.broadcast(ref, 0)
    assert torch.equal(ref, g), "%s: replicas differ" % transport
  assert max(errs) < 1e-6, errs
  for _ in range(20):
    L.check(lib.cpp_ddpg_all_reduce_grads(eng.handle, L.stream_ptr()))
  dp.barrier(); torch.cuda.synchronize()
  n = 500
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(n):
    L.check(lib.cpp_ddpg_all_reduce_grads(eng.handle, L.stream_ptr()))
  b.record(); torch.cuda.synchronize()
  t = torch.tensor([a.elapsed_time(b) / n * 1e3], device="cuda", dtype=torch.float64)
  dp.all_reduce_max(t)
  out[transport] = dict(us_per_call=float(t.item()), max_rel_err=max(errs))
# torch.distributed for reference
g = torch.randn(231667, device="cuda")
for _ in range(20):
  dist.all_reduce(g)
torch.cuda.synchronize(); dp.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(500):
  dist.all_reduce(g)
b.record(); torch.cuda.synchronize()
t = torch.tensor([a.elapsed_time(b) / 500 * 1e3], device="cuda", dtype=torch.float64)
dp.all_reduce_max(t)
out["torch.distributed"] = dict(us_per_call=float(t.item()))
if dp.rank == 0:
  print(json.dumps(dict(world=dp.world_size, floats=int(eng.buffers["grads"].numel()), **out)))
dp.close()
