"""Agent-level engines: own the flat device buffers (torch tensors as containers) and drive the step
functions of libcartpolepp.  One engine per agent; the Network objects of base_network.py are views."""
import os
import ctypes as C
import numpy as np
import torch

from . import _lib

OPTIMISERS = {"GradientDescent": 0, "Momentum": 1, "Adam": 2}


def _require_cuda():
  if not torch.cuda.is_available():
    raise RuntimeError("cartpoleplusplus_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def parse_optimiser(name, args):
  """util.construct_optimiser (util.py:73-76): tf.train.<name>Optimizer(**args) -> (kind, hyper-parameters)"""
  if name not in OPTIMISERS:
    raise ValueError("tf.train.%sOptimizer is not implemented on this path (have %s)" % (name, sorted(OPTIMISERS)))
  a = dict(args)
  hp = dict(lr=float(a.pop("learning_rate")), momentum=0.0, beta1=0.9, beta2=0.999, eps=1e-8)
  if name == "Momentum":
    hp["momentum"] = float(a.pop("momentum"))
  if name == "Adam":
    hp["beta1"] = float(a.pop("beta1", 0.9)); hp["beta2"] = float(a.pop("beta2", 0.999)); hp["eps"] = float(a.pop("epsilon", 1e-8))
  a.pop("use_locking", None); a.pop("name", None)
  if a:
    raise TypeError("unexpected optimiser args %s" % sorted(a))
  return OPTIMISERS[name], hp


class DeviceStager(object):
  """host -> device staging of step inputs through reusable pinned buffers (async on the current stream)"""

  def __init__(self, device):
    self.device = device
    self._pinned = {}
    self._dev = {}
    self._events = {}

  def __call__(self, key, x, dtype=None):
    """returns a device tensor holding x (numpy / torch cpu / torch cuda); dtype None keeps fp16/fp32 as is"""
    if torch.is_tensor(x):
      if x.is_cuda:
        t = x if dtype is None or x.dtype == dtype else x.to(dtype)
        return t.contiguous()
      src = x
    else:
      a = np.asarray(x)
      if a.dtype not in (np.float16, np.float32, np.int32, np.int64):
        a = a.astype(np.float32)
      src = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and src.dtype != dtype:
      src = src.to(dtype)
    src = src.contiguous()
    n, k = src.numel(), (key, src.dtype)
    dev = self._dev.get(k)
    if dev is None or dev.numel() < n:
      dev = torch.empty(max(n, 1), dtype=src.dtype, device=self.device)
      self._dev[k] = dev
    out = dev[:n].view(src.shape)
    if src.is_pinned():
      out.copy_(src, non_blocking=True)
    else:
      pin = self._pinned.get(k)
      if pin is None or pin.numel() < n:
        pin = torch.empty(max(n, 1), dtype=src.dtype, pin_memory=True)
        self._pinned[k] = pin
      # the pinned buffer is reused: wait until the previous async copy out of it has finished
      ev = self._events.get(k)
      if ev is None:
        ev = self._events[k] = torch.cuda.Event()
      else:
        ev.synchronize()
      pin[:n].view(src.shape).copy_(src)
      out.copy_(pin[:n].view(src.shape), non_blocking=True)
      ev.record()
    return out


def state_flag(t):
  if t.dtype == torch.float16:
    return 1
  if t.dtype == torch.float32:
    return 0
  raise TypeError("states must be float16 (replay memory) or float32 (environment), got %s" % t.dtype)


class EngineBase(object):
  def __init__(self):
    _require_cuda()
    self.lib = _lib.lib()
    self.device = torch.device("cuda", torch.cuda.current_device())
    # two staging slots: while a step computes out of one, prefetch() fills the other on a copy stream
    self._stagers = [DeviceStager(self.device), DeviceStager(self.device)]
    self._slot = 0
    self.stage = self._stagers[0]
    self._copy_stream = None
    self._copy_stream2 = None
    self._slot_free = [None, None]       # event: the compute stream is done reading this slot
    self._prefetched = None              # (batch object, staged tensors, copy-done event, slot)
    self.parts = {}           # part name -> (buffer name, offset, size)
    self.buffers = {}
    self.dp, self.lib_comm, self._comm_uid = None, False, None
    self.world_size, self.rank = 1, 0

  # ---- data parallel -----------------------------------------------------------------------------------
  def check_piece_overflow(self, reset=True):
    """raise if an activation exceeded what the fp16 hi + lo piece copy between conv layers can hold (2 x 65504) since the last
    check - the tensor-core route would then have clipped it for the next layer.  Synchronises the device: call it at
    checkpoints / evaluation time, not every step (the CLIs do so once per STATS line)."""
    n = int(self.lib.cpp_piece_overflow_count(1 if reset else 0))
    if n < 0:
      raise _lib.CppError(-2, self.lib.cpp_last_error().decode("utf-8", "replace"))
    if n > 0:
      raise _lib.CppError(-4, "%d conv activations exceeded 131008: the fp16 piece copy of the tensor-core route saturated; "
                              "rerun with cpp_set_option('conv1_tc', 0) (exact fp32 route)" % n)

  def set_data_parallel(self, dp, lib_comm=True, transport=None):
    """make this engine one replica of `dp.world_size`: rank 0's parameters, targets and optimiser state are broadcast (replicas
    must start from identical bits: only gradients are exchanged afterwards), and - on GPUs - the gradient all-reduce moves
    inside the step (captured in its CUDA graph) instead of being issued from here.  transport: "p2p" (default; our own
    all-reduce over NVLink peer memory, csrc/comm.cu; CARTPOLEPP_COMM overrides) or "nccl"."""
    self.dp = dp
    self.world_size, self.rank = dp.world_size, dp.rank
    self.lib_comm, self.transport = False, None
    if not dp.enabled:
      return
    nets = getattr(self, "nets", None) or {}
    if any(getattr(getattr(n, "_final", None), "bn", False) for n in nets.values()):
      # slim.batch_norm statistics of a sharded batch are the shard's, not the global batch's: a per-layer exchange of the
      # channel sums (forward and backward) would be needed to stay equal to the single-replica result - not built
      raise NotImplementedError("--use-batch-norm under data parallelism needs a per-layer statistics all-reduce (not built)")
    for name in ("params", "target_params", "slots", "opt_state"):
      if name in self.buffers:
        dp.broadcast(self.buffers[name], 0)
    import torch.distributed as dist
    if lib_comm and dist.get_backend() == "nccl":
      self.transport = transport or os.environ.get("CARTPOLEPP_COMM", "p2p")
      if self.transport not in ("p2p", "nccl"):
        raise ValueError("transport %r (p2p | nccl)" % self.transport)
      self.lib_comm = True
      self._connect()

  def _connect(self):
    """(re-)join the replica group with the agent object the engine currently holds; collective over all ranks"""
    dp = self.dp
    if self.transport == "nccl":
      _lib.check(self._comm_init(self.handle, dp.rank, dp.world_size, dp.nccl_unique_id()))
    else:
      handle = (C.c_uint8 * 64)()
      _lib.check(self._p2p_prepare(self.handle, dp.rank, dp.world_size, handle))
      _lib.check(self._p2p_connect(self.handle, dp.all_gather_bytes(handle)))
    dp.barrier()

  def _need_global_moments(self, moments, pixels):
    if self.dp is not None and self.dp.enabled and pixels and moments is None:
      raise ValueError("data parallel on pixel states: pass moments=(mean_inv_s1, mean_inv_s2) of the GLOBAL batch "
                       "(ReplayMemory.batch_moments); a shard's own statistics would whiten every replica differently")

  # ---- rollout path (8f row 2) ------------------------------------------------------------------------------------------
  def _action_given(self, fast_fn, exact_fn, states, A):
    """B states -> (B, A) actions through the graph-replayed fast path (tensor-core trunk on an exact fp16 copy of an fp32
    state); a state that is not made of fp16 numbers is re-run through the exact fp32 route.  One pinned read-back."""
    s = self.stage("s_act", states)
    B = int(s.shape[0])
    self._ensure(B)
    n = B * A
    if getattr(self, "_act_dev", None) is None or self._act_dev.numel() < n + 1:
      self._act_dev = torch.zeros(n + 1, dtype=torch.float32, device=self.device)
      self._act_pin = torch.zeros(n + 1, dtype=torch.float32, pin_memory=True)
      self._act_ev = torch.cuda.Event()
    _lib.check(fast_fn(self.handle, _lib.ptr(s), state_flag(s), B, _lib.ptr(self._act_dev), self._stream()))
    self._act_pin[:n + 1].copy_(self._act_dev[:n + 1], non_blocking=True)
    self._act_ev.record()
    self._act_ev.synchronize()
    if self._act_pin[n] != 0:                       # not the env's fp16-valued pixels: keep fp32 exact
      _lib.check(exact_fn(self.handle, _lib.ptr(s), state_flag(s), B, _lib.ptr(self._act_dev), self._stream()))
      self._act_pin[:n].copy_(self._act_dev[:n], non_blocking=True)
      self._act_ev.record()
      self._act_ev.synchronize()
    return self._act_pin[:n].numpy().reshape(B, A).copy()

  def part_view(self, part):
    bufname, off, n = self.parts[part]
    return self.buffers[bufname][off:off + n]

  def _stream(self):
    return _lib.stream_ptr()

  @staticmethod
  def _f(x):
    return C.c_float(float(x))


  # ---- double-buffered input staging ---------------------------------------------------------------------------------
  def prefetch(self, batch):
    """start the host->device copy of `batch` (a Batch of host arrays / pinned tensors) on a copy stream so that it overlaps
    the step that is computing now; the next train_step(batch) called with this very object uses the staged copy"""
    if self._copy_stream is None:
      self._copy_stream = torch.cuda.Stream(device=self.device)
      # CARTPOLEPP_COPY_STREAMS=2: state_2 travels on a second copy stream (two DMA engines share the link)
      self._copy_stream2 = torch.cuda.Stream(device=self.device) if int(os.environ.get("CARTPOLEPP_COPY_STREAMS", "1")) >= 2 else None
    slot = 1 - self._slot
    st = self._stagers[slot]
    cs, cs2 = self._copy_stream, self._copy_stream2
    if self._slot_free[slot] is not None:
      cs.wait_event(self._slot_free[slot])          # a step that read this slot must have finished with it
      if cs2 is not None:
        cs2.wait_event(self._slot_free[slot])
    s2 = None
    if cs2 is not None:
      with torch.cuda.stream(cs2):
        s2 = st("s2", batch.state_2)
        done2 = torch.cuda.Event()
        done2.record(cs2)
    with torch.cuda.stream(cs):
      s1 = st("s1", batch.state_1)
      if s2 is None:
        s2 = st("s2", batch.state_2)
      small = (st("a", batch.action, torch.float32), st("r", batch.reward, torch.float32), st("m", batch.terminal_mask, torch.float32))
      if cs2 is not None:
        cs.wait_event(done2)
      staged = (s1,) + small + (s2,)
      done = torch.cuda.Event()
      done.record(cs)
    self._prefetched = (batch, staged, done, slot)

  def _staged(self, batch):
    """-> (s1, a, r, m, s2) device tensors of `batch`: the prefetched copy when there is one, else staged now"""
    pf, self._prefetched = self._prefetched, None
    if pf is not None and pf[0] is batch:
      torch.cuda.current_stream().wait_event(pf[2])
      self._slot = pf[3]
      self.stage = self._stagers[self._slot]
      return pf[1]
    st = self.stage
    return (st("s1", batch.state_1), st("a", batch.action, torch.float32), st("r", batch.reward, torch.float32),
            st("m", batch.terminal_mask, torch.float32), st("s2", batch.state_2))

  def _release_slot(self):
    """call after the step's work is enqueued: marks when the current staging slot may be overwritten"""
    ev = self._slot_free[self._slot]
    if ev is None:
      ev = self._slot_free[self._slot] = torch.cuda.Event()
    ev.record()
