"""Data parallel on real GPUs (needs >= 2 of them; skipped otherwise): two NCCL ranks train DDPG and NAF on halves of the
same global batch with the global-batch whitening statistics pinned; after every step the replicas must hold IDENTICAL
parameter bits, and they must match a single-GPU run on the whole batch to fp32 summation-order accuracy."""
import os
import socket
import sys
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


MODES = ("p2p", "nccl", "host")


def _free_port():
  s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
  return p


def _moments(batch_states, C_):
  import ctypes as C
  from cartpoleplusplus_b200 import _lib
  lib = _lib.lib()
  x = torch.from_numpy(np.ascontiguousarray(batch_states)).cuda()
  n_pix = x.shape[0] * x.shape[1] * x.shape[2]
  scratch = torch.zeros(int(lib.cpp_moments_scratch_doubles(C_)), dtype=torch.float64, device="cuda")
  out = torch.zeros(2 * C_, dtype=torch.float32, device="cuda")
  _lib.check(lib.cpp_channel_moments(_lib.ptr(x), 1, C.c_int64(n_pix), C_, _lib.ptr(scratch), _lib.ptr(out), _lib.stream_ptr()))
  return out


def _run(kind, world, rank, dp, steps=3, lib_comm=True, transport=None):
  from tests import gpu_util as U
  from oracle import nets_oracle as no
  from oracle.make_golden import ddpg_params, _batch
  shape, Bg = (32, 32, 3, 2, 2), 32
  Cin = 12
  rs = np.random.RandomState(7)
  if kind == "ddpg":
    P = ddpg_params(rs, shape, True)
    nets, eng, o = U.make_ddpg(shape, True, {k: v.numpy() for k, v in P.items()}, batch_size=Bg)
  else:
    P = {}
    for d in (no.naf_value("value", shape, True), no.naf_mu(shape, True), no.naf_l(shape, True)):
      P.update(no.init_params(d, rs))
    P.update(no.retarget({k: v for k, v in P.items() if k.startswith("value/")}, "value", "target_value"))
    naf, nets, eng, o = U.make_naf(shape, True, {k: v.numpy() for k, v in P.items()}, batch_size=Bg, optimiser="Momentum",
                                   optimiser_args={"learning_rate": 0.01, "momentum": 0.9})
  if dp is not None:
    eng.set_data_parallel(dp, lib_comm=lib_comm, transport=transport)
  per = Bg // world
  for step in range(steps):
    batch = _batch(np.random.RandomState(100 + step), Bg, shape)
    m1, m2 = _moments(batch[0], Cin), _moments(batch[4], Cin)
    mine = U.Batch(*[np.ascontiguousarray(x[rank * per:(rank + 1) * per]) for x in batch])
    if kind == "ddpg":
      eng.train_step(mine, moments=(m1, m2))
      if step == 1:
        eng.update_targets()
    else:
      eng.train(mine, moments=(m1, m2))
  torch.cuda.synchronize()
  return torch.cat([eng.buffers["params"], eng.buffers["target_params"]]).clone()


def _worker(rank, world, port, q):
  sys.path.insert(0, ROOT)
  os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.cuda.set_device(rank)
  import torch.distributed as dist
  from cartpoleplusplus_b200 import dp as dpmod
  dp = dpmod.DataParallel(backend="nccl")
  out = {}
  for kind in ("ddpg", "naf"):
    # the all-reduce inside the step's graph (csrc/comm.cu) over NVLink peer memory / through NCCL, and issued from the host
    for mode in MODES:
      mine = _run(kind, world, rank, dp, steps=5, lib_comm=mode != "host", transport=None if mode == "host" else mode)
      gathered = [torch.zeros_like(mine) for _ in range(world)]
      dist.all_gather(gathered, mine)
      assert all(torch.equal(g, gathered[0]) for g in gathered), "%s: replicas diverged (%s)" % (kind, mode)
      out[(kind, mode)] = mine.cpu().numpy()
  dp.barrier()
  if rank == 0:
    q.put(out)
  dp.close()


@pytest.mark.timeout(600)
def test_two_rank_nccl_replicas_identical_and_match_single_gpu():
  if torch.cuda.device_count() < 2:
    pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
  import torch.multiprocessing as mp
  from tests import gpu_util as U
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, 2, _free_port() if r == 0 else 0, q)) for r in range(2)]
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  out = q.get(timeout=500)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  for kind in ("ddpg", "naf"):
    single = _run(kind, 1, 0, None, steps=5).cpu().numpy()
    for mode in MODES:
      U.assert_close(out[(kind, mode)], single, tol=5e-6, what="%s: 2-rank data parallel (%s) vs one GPU on the whole batch" % (kind, mode))
