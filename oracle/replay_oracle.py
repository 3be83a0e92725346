"""Oracle restatement of the reference replay memory.

Follows /root/reference/replay_memory.py:
  * storage layout                :11-38   fp16 state slab of int(buffer_size*load_factor) slots,
                                           int32 state_1_idx/state_2_idx, f32 action/reward/terminal_mask
  * add_episode / _add            :63-118  FIFO free-slot list, evict frees state_1 slot always and
                                           the state_2 slot iff the evicted row was terminal (mask==0)
  * size / random_indexes / batch :120-138 legacy np.random.randint on the global stream, 5 gathers

numpy only, Python 3.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
Pinned: oracle/make_golden.py runs the *real* reference file (print statements converted in
memory, tensorflow/event_log/util stubbed) on the same episode streams and stores its tables
and batches in tests/golden/replay_*.npz; tests/test_oracle_replay.py compares this class
against those, plus the reference's own unit-test expectations
(/root/reference/replay_memory_test.py:19-86).
"""
import collections
import numpy as np

Batch = collections.namedtuple("Batch", "state_1 action reward terminal_mask state_2")


class ReplayOracle:
  def __init__(self, buffer_size, state_shape, action_dim, load_factor=1.5, rng=None):
    assert load_factor >= 1.5, "load_factor has to be at least 1.5"
    self.buffer_size = buffer_size
    self.state_shape = tuple(state_shape)
    self.insert, self.full = 0, False
    self.state_1_idx = np.empty(buffer_size, np.int32)
    self.state_2_idx = np.empty(buffer_size, np.int32)
    self.action = np.empty((buffer_size, action_dim), np.float32)
    self.reward = np.empty((buffer_size, 1), np.float32)
    self.terminal_mask = np.empty((buffer_size, 1), np.float32)
    self.state_buffer_size = int(buffer_size * load_factor)
    self.state = np.empty((self.state_buffer_size,) + self.state_shape, np.float16)
    self.free = collections.deque(range(self.state_buffer_size))
    self.stats = collections.Counter()
    # the reference uses the global np.random stream; tests may inject a RandomState
    self.rng = rng if rng is not None else np.random

  def add_episode(self, initial_state, action_reward_state_sequence):
    self.stats['>add_episode'] += 1
    assert len(action_reward_state_sequence) > 0
    s1 = self.free.popleft()
    self.state[s1] = initial_state
    last = len(action_reward_state_sequence) - 1
    for n, (a, r, s2) in enumerate(action_reward_state_sequence):
      s1 = self._add(s1, a, r, n == last, s2)

  def _add(self, s1_idx, a, r, terminal, s2):
    self.stats['>add'] += 1
    assert 0 <= s1_idx < self.state_buffer_size
    row = self.insert
    if self.full:
      self.free.append(int(self.state_1_idx[row]))
      if self.terminal_mask[row] == 0:
        self.free.append(int(self.state_2_idx[row]))
        self.stats['cache_evicted_s2'] += 1
    self.state_1_idx[row] = s1_idx
    self.action[row] = a
    self.reward[row] = r
    self.terminal_mask[row] = 0.0 if terminal else 1.0
    s2_idx = self.free.popleft()
    self.state_2_idx[row] = s2_idx
    self.state[s2_idx] = s2
    self.insert += 1
    if self.insert >= self.buffer_size:
      self.insert, self.full = 0, True
    return s2_idx

  def size(self):
    return self.buffer_size if self.full else self.insert

  def random_indexes(self, n=1):
    if self.full:
      return self.rng.randint(0, self.buffer_size, n)
    if self.insert == 0:
      return []
    return self.rng.randint(0, self.insert, n)

  def batch_at(self, idxs):
    idxs = np.asarray(idxs, dtype=np.int64)
    return Batch(np.copy(self.state[self.state_1_idx[idxs]]),
                 np.copy(self.action[idxs]),
                 np.copy(self.reward[idxs]),
                 np.copy(self.terminal_mask[idxs]),
                 np.copy(self.state[self.state_2_idx[idxs]]))

  def batch(self, batch_size=None):
    self.stats['>batch'] += 1
    return self.batch_at(self.random_indexes(batch_size))

  def current_stats(self):
    d = dict(self.stats)
    d["free_slots"] = len(self.free)
    return d
