"""GPU-resident drop-in for the reference's ReplayMemory (/root/reference/replay_memory.py:9-163).

Same constructor, methods and slot-allocation order; the storage (fp16 state slab, int32 index tables,
fp32 action/reward/terminal_mask) lives in HBM, the FIFO slot allocator stays on the host (it is
sequential integer bookkeeping), and

  random_indexes  -> cpp_mt_randint  : MT19937 + masked rejection, bit exact with np.random.randint on the
                                       SAME global numpy stream (state is imported/exported around the draw)
  batch           -> cpp_replay_gather : one kernel gathers state_1/state_2 rows and the three small columns

batch() returns a Batch namedtuple of CUDA tensors (fp16 states, fp32 action/reward/terminal_mask with the
reference's shapes); hand it straight to critic.train()/naf.train(), or .cpu().numpy() it.
"""
import collections
import ctypes as C
import sys
import time
import numpy as np
import torch

from . import _lib

Batch = collections.namedtuple("Batch", "state_1 action reward terminal_mask state_2")


class _Sampler(object):
  """a1: bit-exact np.random.randint replacement sharing numpy's global legacy stream"""

  def __init__(self):
    self.lib = _lib.lib()
    self.h = C.c_void_p()
    _lib.check(self.lib.cpp_mt_create(C.byref(self.h)))
    self.private = False

  def __del__(self):
    try:
      self.lib.cpp_mt_destroy(self.h)
    except Exception:
      pass

  def seed(self, seed):
    """switch to a private stream seeded like np.random.seed(seed) (bench / data-parallel replicas)"""
    _lib.check(self.lib.cpp_mt_seed(self.h, C.c_uint32(seed)))
    self.private = True

  def randint(self, high, n):
    out = np.empty(n, dtype=np.int64)
    if not self.private:
      st = np.random.get_state()
      key = np.ascontiguousarray(st[1], dtype=np.uint32)
      _lib.check(self.lib.cpp_mt_set_state(self.h, key.ctypes.data_as(C.c_void_p), C.c_int32(int(st[2]))))
    _lib.check(self.lib.cpp_mt_randint(self.h, C.c_int64(high), C.c_int64(n), out.ctypes.data_as(C.c_void_p)))
    if not self.private:
      pos = C.c_int32()
      _lib.check(self.lib.cpp_mt_get_state(self.h, key.ctypes.data_as(C.c_void_p), C.byref(pos)))
      np.random.set_state((st[0], key, int(pos.value), st[3], st[4]))
    return out


class ReplayMemory(object):
  def __init__(self, buffer_size, state_shape, action_dim, load_factor=1.5, device=None):
    assert load_factor >= 1.5, "load_factor has to be at least 1.5"
    if not torch.cuda.is_available():
      raise RuntimeError("ReplayMemory keeps its storage in HBM: a CUDA device is required (no CPU fallback)")
    self.lib = _lib.lib()
    self.device = device or torch.device("cuda", torch.cuda.current_device())
    self.buffer_size = buffer_size
    self.state_shape = tuple(state_shape)
    self.action_dim = action_dim
    self.insert = 0
    self.full = False
    self.row_elems = int(np.prod(self.state_shape))
    self.pixels = len(self.state_shape) == 5
    self.channels = int(np.prod(self.state_shape[2:])) if self.pixels else 0
    self.n_pix = self.state_shape[0] * self.state_shape[1] if self.pixels else 0

    # host mirrors of the small tables: the slot allocator reads them (replay_memory.py:81-92)
    self.state_1_idx = np.empty(buffer_size, dtype=np.int32)
    self.state_2_idx = np.empty(buffer_size, dtype=np.int32)
    self.terminal_mask = np.empty((buffer_size, 1), dtype=np.float32)
    self.action = np.empty((buffer_size, action_dim), dtype=np.float32)
    self.reward = np.empty((buffer_size, 1), dtype=np.float32)

    self.state_buffer_size = int(buffer_size * load_factor)
    dev = self.device
    self.d_state = torch.empty((self.state_buffer_size, self.row_elems), dtype=torch.float16, device=dev)
    self.d_state_1_idx = torch.zeros(buffer_size, dtype=torch.int32, device=dev)
    self.d_state_2_idx = torch.zeros(buffer_size, dtype=torch.int32, device=dev)
    self.d_action = torch.zeros((buffer_size, action_dim), dtype=torch.float32, device=dev)
    self.d_reward = torch.zeros(buffer_size, dtype=torch.float32, device=dev)
    self.d_mask = torch.zeros(buffer_size, dtype=torch.float32, device=dev)
    # per-slot per-channel (sum, sum of squares) for data-parallel-safe whitening moments
    self.d_slot_stats = (torch.zeros((self.state_buffer_size, 2 * self.channels), dtype=torch.float64, device=dev)
                         if self.pixels else None)

    self.state_free_slots = collections.deque(range(self.state_buffer_size))
    self.stats = collections.Counter()
    self.sampler = _Sampler()

  # ---- insert path (8f row 1): host allocator, one batched H2D + scatter per episode -------------
  def add_episode(self, initial_state, action_reward_state_sequence):
    self.stats['>add_episode'] += 1
    assert len(action_reward_state_sequence) > 0
    n = len(action_reward_state_sequence)
    rows, slots = [], []
    state_1_idx = self.state_free_slots.popleft()           # IndexError when exhausted, like list.pop(0)
    slots.append(state_1_idx)
    for i, (action, reward, _state_2) in enumerate(action_reward_state_sequence):
      rows.append(self.insert)
      state_2_idx = self._add(state_1_idx, action, reward, i == n - 1)
      slots.append(state_2_idx)
      state_1_idx = state_2_idx
    # states -> fp16 (np.float16 storage, replay_memory.py:32) -> HBM
    host = np.empty((n + 1, self.row_elems), dtype=np.float16)
    host[0] = np.asarray(initial_state).reshape(-1)
    for i, (_, _, s2) in enumerate(action_reward_state_sequence):
      host[i + 1] = np.asarray(s2).reshape(-1)
    # an episode longer than the buffer can touch a slot / row twice: the last write wins
    last_of = {}
    for i, sl in enumerate(slots):
      last_of[sl] = i
    slots = sorted(last_of)
    host = host[[last_of[sl] for sl in slots]]
    rows = sorted(set(rows))
    d_slots = torch.tensor(slots, dtype=torch.int64, device=self.device)
    self.d_state.index_copy_(0, d_slots, torch.from_numpy(host).to(self.device))
    d_rows = torch.tensor(rows, dtype=torch.int64, device=self.device)
    r = np.asarray(rows)
    self.d_state_1_idx.index_copy_(0, d_rows, torch.from_numpy(self.state_1_idx[r]).to(self.device))
    self.d_state_2_idx.index_copy_(0, d_rows, torch.from_numpy(self.state_2_idx[r]).to(self.device))
    self.d_action.index_copy_(0, d_rows, torch.from_numpy(self.action[r]).to(self.device))
    self.d_reward.index_copy_(0, d_rows, torch.from_numpy(self.reward[r, 0]).to(self.device))
    self.d_mask.index_copy_(0, d_rows, torch.from_numpy(self.terminal_mask[r, 0]).to(self.device))
    if self.pixels:
      d_slots32 = d_slots.to(torch.int32)
      _lib.check(self.lib.cpp_slot_stats(_lib.ptr(self.d_state), _lib.ptr(d_slots32), C.c_int32(len(slots)),
                                         C.c_int64(self.n_pix), C.c_int32(self.channels),
                                         _lib.ptr(self.d_slot_stats), _lib.stream_ptr()))

  def _add(self, s1_idx, a, r, t):
    self.stats['>add'] += 1
    assert s1_idx >= 0, s1_idx
    assert s1_idx < self.state_buffer_size, s1_idx
    if self.full:
      # always free the state_1 slot of the row about to be clobbered, and its state_2 slot iff that row
      # was terminal (no later row uses it as state_1); replay_memory.py:81-92
      self.state_free_slots.append(int(self.state_1_idx[self.insert]))
      if self.terminal_mask[self.insert] == 0:
        self.state_free_slots.append(int(self.state_2_idx[self.insert]))
        self.stats['cache_evicted_s2'] += 1
    self.state_1_idx[self.insert] = s1_idx
    self.action[self.insert] = a
    self.reward[self.insert] = r
    self.terminal_mask[self.insert] = 0.0 if t else 1.0
    s2_idx = self.state_free_slots.popleft()
    self.state_2_idx[self.insert] = s2_idx
    self.insert += 1
    if self.insert >= self.buffer_size:
      self.insert = 0
      self.full = True
    return s2_idx

  def reset_from_event_log(self, log_file):
    """replay_memory.py:40-61: prefill from a recorded event log until the memory is full"""
    import sys
    import time
    from . import event_log
    elr = event_log.EventLogReader(log_file)
    num_episodes = num_events = 0
    start = time.time()
    for episode in elr.entries():
      initial_state = None
      action_reward_state_sequence = []
      for event_id, event in enumerate(episode.event):
        if event_id == 0:
          assert len(event.action) == 0
          assert not event.HasField("reward")
          initial_state = event_log.read_state_from_event(event)
        else:
          action_reward_state_sequence.append((np.asarray(event.action, dtype=np.float32), event.reward,
                                               event_log.read_state_from_event(event)))
        num_events += 1
      num_episodes += 1
      self.add_episode(initial_state, action_reward_state_sequence)
      if self.full:
        break
    sys.stderr.write("reset_from_event_log \"%s\" num_episodes=%d num_events=%d took %s sec\n"
                     % (log_file, num_episodes, num_events, time.time() - start))

  # ---- sample path --------------------------------------------------------------------------------
  def size(self):
    return self.buffer_size if self.full else self.insert

  def random_indexes(self, n=1):
    if self.full:
      return self.sampler.randint(self.buffer_size, n)
    elif self.insert == 0:  # empty
      return []
    else:
      return self.sampler.randint(self.insert, n)

  def batch_at(self, idxs, d_idxs=None, out=None):
    """gather the rows `idxs` (host int64 array) -> Batch of CUDA tensors.  d_idxs: the same indexes already on
    the device; out: a Batch of preallocated tensors to fill (steady-state loops avoid allocator traffic)"""
    B = len(idxs)
    dev = self.device
    if out is not None:
      s1, a, r, m, s2 = out
      assert s1.shape[0] == B
    else:
      s1 = torch.empty((B,) + self.state_shape, dtype=torch.float16, device=dev)
      s2 = torch.empty((B,) + self.state_shape, dtype=torch.float16, device=dev)
      a = torch.empty((B, self.action_dim), dtype=torch.float32, device=dev)
      r = torch.empty((B, 1), dtype=torch.float32, device=dev)
      m = torch.empty((B, 1), dtype=torch.float32, device=dev)
    if B == 0:
      return Batch(s1, a, r, m, s2)
    if d_idxs is None:
      d_idxs = torch.from_numpy(np.ascontiguousarray(idxs, dtype=np.int64)).to(dev)
    _lib.check(self.lib.cpp_replay_gather(_lib.ptr(self.d_state), _lib.ptr(self.d_state_1_idx), _lib.ptr(self.d_state_2_idx),
                                          _lib.ptr(self.d_action), _lib.ptr(self.d_reward), _lib.ptr(self.d_mask),
                                          _lib.ptr(d_idxs), C.c_int32(B), C.c_int64(self.row_elems), C.c_int32(self.action_dim),
                                          _lib.ptr(s1), _lib.ptr(s2), _lib.ptr(a), _lib.ptr(r), _lib.ptr(m), _lib.stream_ptr()))
    return Batch(s1, a, r, m, s2)

  def batch(self, batch_size=None):
    self.stats['>batch'] += 1
    idxs = self.random_indexes(batch_size)
    return self.batch_at(idxs)

  def batch_moments(self, d_idxs, which, out=None):
    """whitening (mean, rsqrt(var+1e-6)) of the GLOBAL batch `d_idxs` from the per-slot sums; which=1|2"""
    assert self.pixels
    table = self.d_state_1_idx if which == 1 else self.d_state_2_idx
    if out is None:
      out = torch.empty(2 * self.channels, dtype=torch.float32, device=self.device)
    _lib.check(self.lib.cpp_moments_from_slots(_lib.ptr(self.d_slot_stats), _lib.ptr(table), _lib.ptr(d_idxs),
                                               C.c_int32(int(d_idxs.numel())), C.c_int64(self.n_pix), C.c_int32(self.channels),
                                               _lib.ptr(out), _lib.stream_ptr()))
    return out

  def current_stats(self):
    current_stats = dict(self.stats)
    current_stats["free_slots"] = len(self.state_free_slots)
    return current_stats
