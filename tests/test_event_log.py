"""Event-log format: the hand-written descriptor must parse bytes laid out by the reference's
event.proto (field numbers and wire types pinned by a hand-encoded episode), the framing is struct '=l' + message
(event_log.py:50-58,103-111), PNG renders round-trip at 8 bits, and - on a GPU - ReplayMemory.reset_from_event_log
(replay_memory.py:40-61) fills the device-resident memory exactly like add_episode on the same episodes."""
import os
import struct
import numpy as np
import pytest

from cartpoleplusplus_b200 import event_log as el


def _tag(field, wire):
  return bytes([(field << 3) | wire])


def _varint(n):
  out = b""
  while True:
    b, n = n & 0x7f, n >> 7
    out += bytes([b | (0x80 if n else 0)])
    if not n:
      return out


def _len_delim(field, payload):
  return _tag(field, 2) + _varint(len(payload)) + payload


def _floats(field, vals):       # proto2 repeated float, not packed: one fixed32 per element
  return b"".join(_tag(field, 5) + struct.pack("<f", v) for v in vals)


def test_hand_encoded_episode_parses(tmp_path):
  cart, pole = [1, 2, 3, 4, 5, 6, 7], [8, 9, 10, 11, 12, 13, 14]
  state = _floats(1, cart) + _floats(2, pole)                          # State.cart_pose = 1, pole_pose = 2
  ev0 = _len_delim(2, state) + _len_delim(2, state)                     # Event.state = 2 (two action repeats), no action / reward
  ev1 = _floats(1, [0.5, -0.25]) + _len_delim(2, state) + _len_delim(2, state) + _tag(3, 5) + struct.pack("<f", 1.5)
  episode = _len_delim(1, ev0) + _len_delim(1, ev1)                     # Episode.event = 1
  p = str(tmp_path / "log")
  with open(p, "wb") as f:
    f.write(struct.pack("=l", len(episode)) + episode)
  eps = list(el.EventLogReader(p).entries())
  assert len(eps) == 1 and len(eps[0].event) == 2
  e0, e1 = eps[0].event
  assert len(e0.action) == 0 and not e0.HasField("reward")
  assert list(e1.action) == [0.5, -0.25] and e1.reward == 1.5
  s = el.read_state_from_event(e1)
  assert s.shape == (2, 2, 7) and np.array_equal(s[1][0], cart) and np.array_equal(s[0][1], pole)


def _episodes(rs, pixels, n_eps):
  out = []
  for _ in range(n_eps):
    T = rs.randint(2, 6)
    if pixels:
      mk = lambda: rs.randint(0, 256, (12, 10, 3, 2, 3)).astype(np.float32) / np.float32(255)
    else:
      mk = lambda: rs.randn(2, 2, 7).astype(np.float32)
    out.append((mk(), [(rs.uniform(-1, 1, (1, 2)).astype(np.float32), float(rs.rand()), mk()) for _ in range(T)]))
  return out


@pytest.mark.parametrize("pixels", [False, True])
def test_write_read_round_trip(tmp_path, pixels):
  rs = np.random.RandomState(3)
  eps = _episodes(rs, pixels, 3)
  p = str(tmp_path / "log")
  log = el.EventLog(p, pixels)
  for init, seq in eps:
    log.reset()
    log.add_just_state(init)
    for a, r, s2 in seq:
      log.add(s2, a, r)
  log.close()
  got = list(el.EventLogReader(p).entries())
  assert len(got) == len(eps)
  for ep, (init, seq) in zip(got, eps):
    assert len(ep.event) == len(seq) + 1
    tol = 0 if pixels else 0          # 8-bit renders and float32 poses both survive exactly
    assert np.abs(el.read_state_from_event(ep.event[0]) - init).max() <= tol
    for ev, (a, r, s2) in zip(list(ep.event)[1:], seq):
      assert np.allclose(np.asarray(ev.action, dtype=np.float32), a[0]) and abs(ev.reward - np.float32(r)) < 1e-7
      assert np.abs(el.read_state_from_event(ev) - s2).max() <= tol


@pytest.mark.gpu
@pytest.mark.parametrize("pixels", [False, True])
def test_reset_from_event_log_equals_add_episode(tmp_path, pixels):
  import torch
  from cartpoleplusplus_b200.replay_memory import ReplayMemory
  rs = np.random.RandomState(5)
  eps = _episodes(rs, pixels, 6)
  p = str(tmp_path / "log")
  log = el.EventLog(p, pixels)
  for init, seq in eps:
    log.reset(); log.add_just_state(init)
    for a, r, s2 in seq:
      log.add(s2, a, r)
  log.close()
  shape = (12, 10, 3, 2, 3) if pixels else (2, 2, 7)
  a_mem, b_mem = ReplayMemory(64, shape, 2), ReplayMemory(64, shape, 2)
  a_mem.reset_from_event_log(p)
  for init, seq in eps:
    b_mem.add_episode(init, seq)
  assert a_mem.size() == b_mem.size() and a_mem.insert == b_mem.insert
  n = a_mem.size()
  for name in ("state_1_idx", "state_2_idx", "action", "reward", "terminal_mask"):
    assert np.array_equal(getattr(a_mem, name)[:n], getattr(b_mem, name)[:n]), name
  assert torch.equal(a_mem.d_state, b_mem.d_state) or torch.equal(a_mem.d_state[:n + 8], b_mem.d_state[:n + 8])
  np.random.seed(0); ba = a_mem.batch(16)
  np.random.seed(0); bb = b_mem.batch(16)
  for x, y in zip(ba, bb):
    assert torch.equal(x, y)


def test_synthetic_env_writes_an_event_log_the_reader_understands(tmp_path):
  """--event-log-out (bullet_cartpole.py:90-94,221-222,283-285): episodes land in the reference's framed format"""
  import argparse
  from cartpoleplusplus_b200 import synthetic_env
  p = argparse.ArgumentParser()
  synthetic_env.add_opts(p)
  path = str(tmp_path / "events.log")
  o = p.parse_args(["--event-log-out", path, "--max-episode-len", "4", "--action-repeats", "2"])
  env = synthetic_env.SyntheticCartpole(o, discrete_actions=False)
  for _ in range(2):
    env.reset()
    done = False
    while not done:
      _, _, done, _ = env.step(np.array([[0.25, -0.5]], dtype=np.float32))
  env.reset()                                       # flushes the second episode, as in the reference
  episodes = list(el.EventLogReader(path).entries())
  assert len(episodes) == 2
  for ep in episodes:
    assert len(ep.event) >= 2 and len(ep.event[0].action) == 0
    assert list(ep.event[1].action) == [0.25, -0.5] and ep.event[1].reward == 1.0
    assert el.read_state_from_event(ep.event[1]).shape == (2, 2, 7)
