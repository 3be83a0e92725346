"""Data-parallel host logic at world_size 2 on CPU (gloo): identical MT19937 index streams on every rank
(through the C ABI sampler, a1), rank slicing, ONE sum all-reduce over the flat [actor | critic | loss] buffer, and
redundant clip + apply keeping replicas bit-identical.  The per-shard gradients come from the oracle (no GPU here); the
GPU side of the same algebra is tests/test_gpu_nets.py::test_ddpg_data_parallel_linearity."""
import os
import socket
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
  return p


def _worker(rank, world, port, B):
  sys.path.insert(0, ROOT)
  os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.set_num_threads(1)
  from cartpoleplusplus_b200 import dp as dpmod
  from cartpoleplusplus_b200.replay_memory import _Sampler
  from oracle import nets_oracle as no
  from oracle.make_golden import ddpg_params, _batch
  dp = dpmod.DataParallel(backend="gloo")
  assert dp.enabled and dp.world_size == world and dp.rank == rank
  # a1: every replica draws the same global index vector
  smp = _Sampler(); smp.seed(0)
  N = 500
  idxs = smp.randint(N, B * world)
  ref = np.random.RandomState(0).randint(0, N, B * world)
  assert np.array_equal(idxs, ref)
  gathered = [torch.zeros(B * world, dtype=torch.int64) for _ in range(world)]
  dist.all_gather(gathered, torch.from_numpy(idxs))
  assert all(torch.equal(g, gathered[0]) for g in gathered)
  mine = dp.shard(idxs, B)
  assert np.array_equal(mine, idxs[rank * B:(rank + 1) * B])
  # a "replay" every rank holds in full (seeded identically), low-dim DDPG
  shape = (2, 2, 7)
  rs = np.random.RandomState(3)
  P = ddpg_params(rs, shape, False)
  table = _batch(rs, N, shape)                                   # (s1, a, r, mask, s2) rows of the whole memory
  def rows(ix):
    return tuple(np.ascontiguousarray(t[ix]) for t in table)
  orc = no.DDPGOracle(shape, False, {k: v.clone() for k, v in P.items()})
  Bg = B * world
  # ---- backward on this rank's shard: actor gradient is a batch SUM, critic gradient / loss are means -> rescale to 1/B_global
  sb = rows(mine)
  ga, _, _, _ = no.ddpg_actor_grads(orc.actor, orc.critic, orc.P, sb[0])
  gc, loss, _, _ = no.ddpg_critic_grads(orc.critic, orc.tactor, orc.tcritic, orc.P, sb, orc.discount)
  flat = torch.cat([g.reshape(-1) for g in ga] + [g.reshape(-1) * (B / Bg) for g in gc] + [loss.reshape(1) * (B / Bg)])
  dp.all_reduce_sum(flat)                                        # the single collective of the step
  # ---- the same quantities on the full global batch
  fb = rows(idxs)
  ga_f, _, _, _ = no.ddpg_actor_grads(orc.actor, orc.critic, orc.P, fb[0])
  gc_f, loss_f, _, _ = no.ddpg_critic_grads(orc.critic, orc.tactor, orc.tcritic, orc.P, fb, orc.discount)
  full = torch.cat([g.reshape(-1) for g in ga_f] + [g.reshape(-1) for g in gc_f] + [loss_f.reshape(1)])
  err = float((flat - full).abs().max() / full.abs().max())
  assert err < 1e-12, err
  # ---- redundant clip + apply: replicas stay bit-identical
  na = sum(g.numel() for g in ga)
  new = []
  for part, lr in ((flat[:na], orc.actor_lr), (flat[na:-1], orc.critic_lr)):
    norm = torch.sqrt((part ** 2).sum())
    new.append(part * (orc.clip * torch.minimum(1.0 / norm, torch.tensor(1.0 / orc.clip, dtype=part.dtype))) * lr)
  upd = torch.cat(new)
  got = [torch.zeros_like(upd) for _ in range(world)]
  dist.all_gather(got, upd)
  assert all(torch.equal(g, got[0]) for g in got)
  t = torch.tensor([1.0 + rank], dtype=torch.float64)
  dp.all_reduce_max(t)
  assert float(t) == float(world)
  dp.barrier()
  dp.close()


@pytest.mark.timeout(300)
def test_data_parallel_world2_gloo():
  mp.spawn(_worker, args=(2, _free_port(), 16), nprocs=2, join=True)
