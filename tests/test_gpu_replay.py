"""GPU parity: replay memory (a1, a2, 8f row 1) through the drop-in class / C ABI, bit exact against the
outputs of the real reference replay_memory.py (tests/golden/replay_*.npz)."""
import json
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def episodes_from(g):
  e = 0
  while "ep%d_init" % e in g:
    acts, rews, sts = g["ep%d_actions" % e], g["ep%d_rewards" % e], g["ep%d_states" % e]
    yield g["ep%d_init" % e], [(acts[i], float(rews[i]), sts[i]) for i in range(len(rews))]
    e += 1


@pytest.mark.parametrize("name", ["small", "ragged", "pixels"])
def test_matches_reference_run_bit_exact(golden_dir, name):
  from cartpoleplusplus_b200.replay_memory import ReplayMemory
  g = np.load(os.path.join(golden_dir, "replay_%s.npz" % name))
  c = json.loads(str(g["meta"]))
  rm = ReplayMemory(c["buffer_size"], c["state_shape"], 2, c["load_factor"])
  np.random.seed(1000 + c["seed"])                # the reference consumed the global numpy stream
  for e, (init, seq) in enumerate(episodes_from(g)):
    rm.add_episode(init, seq)
    b = rm.batch(c["batch"])
    ins, full, size, nfree = g["e%d_insert_full_size" % e]
    assert (rm.insert, int(rm.full), rm.size(), len(rm.state_free_slots)) == (ins, full, size, nfree)
    assert np.array_equal(rm.state_1_idx[:size], g["e%d_s1idx" % e])
    assert np.array_equal(rm.state_2_idx[:size], g["e%d_s2idx" % e])
    assert np.array_equal(np.array(rm.state_free_slots, dtype=np.int32), g["e%d_free" % e])
    assert np.array_equal(rm.d_state_1_idx[:size].cpu().numpy(), g["e%d_s1idx" % e])
    for f, v in zip(b._fields, b):
      ref = g["e%d_batch_%s" % (e, f)]
      got = v.cpu().numpy()
      assert got.dtype == ref.dtype and got.shape == ref.shape and np.array_equal(got, ref), (e, f)


def test_reference_unit_test_expectations():
  """/root/reference/replay_memory_test.py:19-86"""
  from cartpoleplusplus_b200.replay_memory import ReplayMemory
  rm = ReplayMemory(buffer_size=3, state_shape=(2, 3), action_dim=2, load_factor=2)
  assert rm.size() == 0 and rm.random_indexes() == []
  b = rm.batch(4)
  assert len(b) == 5 and all(len(x) == 0 for x in b)
  assert rm.insert == 0 and rm.full is False
  rm.add_episode([[11, 12, 13], [14, 15, 16]],
                 [(17, 18, [[21, 22, 23], [24, 25, 26]]), (27, 28, [[31, 32, 33], [34, 35, 36]]), (37, 38, [[41, 42, 43], [44, 45, 46]])])
  assert rm.size() == 3 and rm.insert == 0 and rm.full is True
  idxs = rm.random_indexes(n=100)
  assert len(idxs) == 100 and sorted(set(idxs.tolist())) == [0, 1, 2]
  state = rm.d_state.cpu().numpy().reshape(-1, 2, 3)
  assert [state[i][0][0] for i in range(4)] == [11, 21, 31, 41]
  rm = ReplayMemory(buffer_size=3, state_shape=(2, 3), action_dim=2, load_factor=2)
  def s_for(i):
    return (np.array(range(1, 7)) + (10 * i)).reshape(2, 3)
  rm.add_episode(s_for(0), [((i * 10) + 7, (i * 10) + 8, s_for(i)) for i in range(1, 5)])
  rm.add_episode(s_for(5), [((i * 10) + 7, (i * 10) + 8, s_for(i)) for i in range(6, 9)])
  assert rm.size() == 3
  b = rm.batch_at(np.array([0, 1, 2]))
  assert np.array_equal(b.reward.cpu().numpy(), [[88], [68], [78]])
  assert np.array_equal(b.terminal_mask.cpu().numpy(), [[0], [1], [1]])
  with pytest.raises(AssertionError):
    ReplayMemory(3, (2, 3), 2, load_factor=1.2)


def test_soak_invariant():
  """consistency property of /root/reference/replay_memory.py:166-200 (bounded run)"""
  from cartpoleplusplus_b200.replay_memory import ReplayMemory
  rm = ReplayMemory(buffer_size=43, state_shape=(2, 3), action_dim=2)
  rs = np.random.RandomState(6)
  np.random.seed(5)
  def s(i):
    i = (i * 10) % 199
    return [[i + 1, 0, 0], [0, 0, 0]]
  terminals, i = set(), 0
  for _ in range(120):
    init, seq = s(i), []
    for _ in range(int(3 + rs.rand() * 5)):
      i += 1
      seq.append(((i, 0), i, s(i)))
    rm.add_episode(init, seq)
    terminals.add(i)
    for _ in range(3):
      b = [x.cpu().numpy() for x in rm.batch(13)]
      for j in range(13):
        r = int(b[2][j][0])
        assert b[0][j][0][0] == (((r - 1) * 10) % 199) + 1
        assert b[1][j][0] == r
        assert b[3][j] == (0 if r in terminals else 1)
        assert b[4][j][0][0] == ((r * 10) % 199) + 1
    i += 1


def test_random_indexes_consumes_numpy_global_stream():
  from cartpoleplusplus_b200.replay_memory import ReplayMemory
  rm = ReplayMemory(buffer_size=50, state_shape=(2, 3), action_dim=2)
  rm.add_episode(np.zeros((2, 3)), [((0, 0), 1.0, np.zeros((2, 3)))] * 30)
  np.random.seed(123)
  a = rm.random_indexes(17); x = np.random.randn(3); b = rm.random_indexes(256)
  np.random.seed(123)
  a2 = np.random.randint(0, 30, 17); x2 = np.random.randn(3); b2 = np.random.randint(0, 30, 256)
  assert np.array_equal(a, a2) and np.array_equal(x, x2) and np.array_equal(b, b2)


def test_full_size_gather_round_trip_and_slot_moments():
  """c3-sized rows (64x64x9 fp16): the gathered rows are byte-identical to the slab rows, and the whitening
  moments assembled from per-slot sums equal the moments of the gathered batch"""
  import ctypes as C
  from cartpoleplusplus_b200 import _lib
  from cartpoleplusplus_b200.replay_memory import ReplayMemory
  shape = (64, 64, 3, 1, 3)
  rm = ReplayMemory(buffer_size=300, state_shape=shape, action_dim=2)
  rs = np.random.RandomState(3)
  for _ in range(12):
    L = 30
    st = lambda: (rs.randint(0, 256, shape).astype(np.float16) / np.float16(255))
    rm.add_episode(st(), [(rs.uniform(-1, 1, (1, 2)), 1.0, st()) for _ in range(L)])
  np.random.seed(0)
  idxs = rm.random_indexes(256)
  b = rm.batch_at(idxs)
  slab = rm.d_state
  want1 = slab[torch.from_numpy(rm.state_1_idx[idxs].astype(np.int64)).cuda()].view(b.state_1.shape)
  want2 = slab[torch.from_numpy(rm.state_2_idx[idxs].astype(np.int64)).cuda()].view(b.state_2.shape)
  assert torch.equal(b.state_1, want1) and torch.equal(b.state_2, want2)
  assert np.array_equal(b.action.cpu().numpy(), rm.action[idxs])
  d_idxs = torch.from_numpy(idxs).cuda()
  for which, batch_states in ((1, b.state_1), (2, b.state_2)):
    mi = rm.batch_moments(d_idxs, which).cpu().numpy().astype(np.float64)
    x = batch_states.cpu().numpy().astype(np.float64).reshape(-1, 9)
    mean = x.mean(0); inv = 1.0 / np.sqrt(x.var(0) + 1e-6)
    np.testing.assert_allclose(mi[:9], mean, rtol=2e-7)
    np.testing.assert_allclose(mi[9:], inv, rtol=2e-7)
    lib = _lib.lib()
    scratch = torch.zeros(int(lib.cpp_moments_scratch_doubles(9)), dtype=torch.float64, device="cuda")
    out = torch.zeros(18, dtype=torch.float32, device="cuda")
    _lib.check(lib.cpp_channel_moments(_lib.ptr(batch_states), 1, C.c_int64(256 * 64 * 64), 9, _lib.ptr(scratch), _lib.ptr(out), _lib.stream_ptr()))
    np.testing.assert_allclose(out.cpu().numpy(), mi, rtol=2e-7)
