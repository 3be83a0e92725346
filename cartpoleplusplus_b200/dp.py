"""Data-parallel plumbing: one process per GPU, torch.distributed for rendezvous and the
single gradient all-reduce (NCCL over NVLink/NVSwitch; gloo in CPU tests).

Sharding contract: every rank holds a full replica (parameters, targets, optimiser slots, replay memory) and
draws the SAME global index vector from identically seeded MT19937 streams; rank r trains on
idxs[r*B : (r+1)*B].  Critic/NAF gradients and the loss are pre-scaled by 1/B_global and the actor gradient is a
batch sum, so the collective is a plain SUM over the flat gradient buffer; clip + optimiser + target update then
run redundantly (and deterministically) on every rank, keeping replicas bit-identical.  Whitening statistics of
the global batch come from per-slot sums (ReplayMemory.batch_moments), so no second collective is needed."""
import os
import torch
import torch.distributed as dist


class DataParallel(object):
  def __init__(self, backend=None):
    self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
    self.rank = int(os.environ.get("RANK", "0"))
    self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    self.enabled = self.world_size > 1
    if self.enabled and not dist.is_initialized():
      os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
      os.environ.setdefault("MASTER_PORT", "29500")
      if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
      if backend == "nccl":
        torch.cuda.set_device(self.local_rank)
      dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world_size)

  def shard(self, idxs, per_rank):
    """this rank's slice of the global index vector"""
    assert len(idxs) == per_rank * self.world_size
    return idxs[self.rank * per_rank:(self.rank + 1) * per_rank]

  def broadcast(self, t, src=0):
    """replicas start (and resume) from rank `src`'s bits: parameters, targets, optimiser slots"""
    if self.enabled:
      dist.broadcast(t, src=src)
    return t

  def nccl_unique_id(self):
    """ncclGetUniqueId on rank 0, handed to every rank: the library creates its OWN communicator from it (csrc/comm.cu),
    torch.distributed is only the rendezvous"""
    import ctypes as C
    from . import _lib
    buf = (C.c_uint8 * 128)()
    if self.rank == 0:
      _lib.check(_lib.lib().cpp_nccl_unique_id(buf))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    out = (C.c_uint8 * 128)(*t.cpu().tolist())
    return out

  def all_gather_bytes(self, buf):
    """every rank's `buf` (ctypes uint8 array of one fixed size) -> one ctypes array, rank-major (CUDA IPC handles)"""
    import ctypes as C
    n = len(buf)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
    out = [torch.zeros(n, dtype=torch.uint8, device=dev) for _ in range(self.world_size)]
    dist.all_gather(out, mine)
    flat = torch.cat(out).cpu().tolist()
    return (C.c_uint8 * len(flat))(*flat)

  def all_reduce_sum(self, flat):
    if self.enabled:
      dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat

  def all_reduce_max(self, t):
    if self.enabled:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t

  def barrier(self):
    if self.enabled:
      dist.barrier()

  def close(self):
    if self.enabled and dist.is_initialized():
      dist.destroy_process_group()
