"""The oracle's replay restatement vs outputs of the REAL reference replay_memory.py
(tests/golden/replay_*.npz, made by oracle/make_golden.py) and vs the reference's own unit-test
expectations (/root/reference/replay_memory_test.py:19-86)."""
import json
import os
import numpy as np
import pytest

from oracle.replay_oracle import ReplayOracle
from oracle.mt19937_oracle import MT19937


def episodes_from(g):
  e = 0
  while "ep%d_init" % e in g:
    acts, rews, sts = g["ep%d_actions" % e], g["ep%d_rewards" % e], g["ep%d_states" % e]
    yield g["ep%d_init" % e], [(acts[i], float(rews[i]), sts[i]) for i in range(len(rews))]
    e += 1


@pytest.mark.parametrize("name", ["small", "ragged", "pixels"])
def test_matches_reference_run(golden_dir, name):
  g = np.load(os.path.join(golden_dir, "replay_%s.npz" % name))
  c = json.loads(str(g["meta"]))
  rs = np.random.RandomState(1000 + c["seed"])        # the reference ran on the seeded global stream
  rm = ReplayOracle(c["buffer_size"], c["state_shape"], 2, c["load_factor"], rng=rs)
  for e, (init, seq) in enumerate(episodes_from(g)):
    rm.add_episode(init, seq)
    b = rm.batch(c["batch"])
    ins, full, size, nfree = g["e%d_insert_full_size" % e]
    assert (rm.insert, int(rm.full), rm.size(), len(rm.free)) == (ins, full, size, nfree)
    assert np.array_equal(rm.state_1_idx[:size], g["e%d_s1idx" % e])
    assert np.array_equal(rm.state_2_idx[:size], g["e%d_s2idx" % e])
    assert np.array_equal(np.array(rm.free, dtype=np.int32), g["e%d_free" % e])
    for f, v in zip(b._fields, b):
      ref = g["e%d_batch_%s" % (e, f)]
      assert v.dtype == ref.dtype and np.array_equal(v, ref), (e, f)
  n = rm.size()
  assert np.array_equal(rm.action[:n], g["final_action"])
  assert np.array_equal(rm.reward[:n], g["final_reward"])
  assert np.array_equal(rm.terminal_mask[:n], g["final_mask"])


def test_reference_unit_test_expectations():
  # replay_memory_test.py:19-30
  rm = ReplayOracle(buffer_size=3, state_shape=(2, 3), action_dim=2, load_factor=2)
  assert rm.size() == 0 and rm.random_indexes() == []
  b = rm.batch(4)
  assert len(b) == 5 and all(len(x) == 0 for x in b)
  assert rm.insert == 0 and rm.full is False
  # :32-56
  rm.add_episode([[11, 12, 13], [14, 15, 16]],
                 [(17, 18, [[21, 22, 23], [24, 25, 26]]), (27, 28, [[31, 32, 33], [34, 35, 36]]),
                  (37, 38, [[41, 42, 43], [44, 45, 46]])])
  assert rm.size() == 3 and rm.insert == 0 and rm.full is True
  idxs = rm.random_indexes(n=100)
  assert len(idxs) == 100 and sorted(set(idxs)) == [0, 1, 2]
  assert [rm.state[i][0][0] for i in range(4)] == [11, 21, 31, 41]
  # :58-86
  rm = ReplayOracle(buffer_size=3, state_shape=(2, 3), action_dim=2, load_factor=2)
  def s_for(i):
    return (np.array(range(1, 7)) + (10 * i)).reshape(2, 3)
  rm.add_episode(s_for(0), [((i * 10) + 7, (i * 10) + 8, s_for(i)) for i in range(1, 5)])
  rm.add_episode(s_for(5), [((i * 10) + 7, (i * 10) + 8, s_for(i)) for i in range(6, 9)])
  assert rm.size() == 3
  b = rm.batch_at([0, 1, 2])
  assert np.array_equal(b.reward, [[88], [68], [78]])
  assert np.array_equal(b.terminal_mask, [[0], [1], [1]])


def test_soak_invariant():
  """the consistency property of /root/reference/replay_memory.py:166-200, bounded"""
  rm = ReplayOracle(buffer_size=43, state_shape=(2, 3), action_dim=2, rng=np.random.RandomState(5))
  rs = np.random.RandomState(6)
  def s(i):
    i = (i * 10) % 199
    return [[i + 1, 0, 0], [0, 0, 0]]
  terminals, i = set(), 0
  for _ in range(400):
    init, seq = s(i), []
    for _ in range(int(3 + rs.rand() * 5)):
      i += 1
      seq.append(((i, 0), i, s(i)))
    rm.add_episode(init, seq)
    terminals.add(i)
    for _ in range(7):
      b = rm.batch(13)
      for j in range(13):
        r = int(b.reward[j][0])
        assert b.state_1[j][0][0] == (((r - 1) * 10) % 199) + 1
        assert b.action[j][0] == r
        assert b.terminal_mask[j] == (0 if r in terminals else 1)
        assert b.state_2[j][0][0] == ((r * 10) % 199) + 1
    i += 1


def test_mt19937_restatement_vs_numpy_kat(golden_dir):
  k = json.load(open(os.path.join(golden_dir, "mt19937_kat.json")))
  for case in k["kat"]:
    m = MT19937(case["seed"])
    for out in case["outs"]:
      assert m.randint(case["high"], case["n"]).tolist() == out, case["seed"]
  # Appendix F first-five vectors
  assert MT19937(0).randint(22000, 5).tolist() == [2732, 10799, 9845, 19648, 13123]
  assert MT19937(42).randint(22000, 5).tolist() == [15795, 860, 5390, 21575, 11964]
  assert MT19937(7).randint(3, 5).tolist() == [0, 1, 2, 0, 1]


def test_mt19937_state_roundtrip_with_numpy():
  rs = np.random.RandomState(77)
  rs.randn(3)
  m = MT19937()
  m.set_numpy_state(rs.get_state())
  assert m.randint(22000, 300).tolist() == rs.randint(0, 22000, 300).tolist()
