"""SaverUtil (reference util.py:88-131): checkpoint save / restore of the device-resident variables.

CPU part: file layout, the `checkpoint` index, shape checks - against a stand-in engine that keeps its flat buffers on the
host.  GPU part: a resumed run reproduces the uninterrupted one bit for bit (parameters, targets, Adam slots and powers)."""
import collections
import os
import numpy as np
import pytest
import torch

from cartpoleplusplus_b200 import base_network, util


class _HostNet(base_network.Network):
  def __init__(self, namespace, shapes):
    base_network.Network.__init__(self, namespace)
    self._shapes = shapes

  def _variables(self):
    out, off = [], 0
    for name, shape in self._shapes:
      n = int(np.prod(shape))
      out.append(base_network.Var("%s/%s" % (self.namespace, name), tuple(shape), off, n))
      off += n
    return out


class _HostEngine(object):
  """the three attributes SaverUtil uses: nets, buffers, part_view"""

  def __init__(self, seed):
    rs = np.random.RandomState(seed)
    shapes = [("h0/weights", (4, 3)), ("h0/biases", (3,)), ("out/weights", (3, 2)), ("out/biases", (2,))]
    self.nets = collections.OrderedDict([("model", _HostNet("model", shapes)), ("target", _HostNet("target_model", shapes))])
    n = 4 * 3 + 3 + 3 * 2 + 2
    self.buffers = dict(params=torch.from_numpy(rs.randn(2 * n).astype(np.float32)),
                        slots=torch.from_numpy(rs.randn(2 * n).astype(np.float32)),
                        opt_state=torch.from_numpy(rs.rand(4).astype(np.float32)))
    self.parts = dict(model=("params", 0, n), target=("params", n, n))
    for part, net in self.nets.items():
      net._engine, net._part = self, part

  def part_view(self, part):
    b, off, n = self.parts[part]
    return self.buffers[b][off:off + n]


def test_saver_file_layout_and_round_trip(tmp_path):
  d = str(tmp_path / "ckpts")
  e = _HostEngine(1)
  want = {k: v.clone() for k, v in e.buffers.items()}
  s = util.SaverUtil(e, d, save_freq=3600)                       # nothing there yet: saves the initial variables
  index = open(os.path.join(d, "checkpoint")).read().splitlines()
  assert index[0].startswith('model_checkpoint_path: "ckpt.')
  first = index[0].split('"')[1]
  with np.load(os.path.join(d, first + ".npz")) as z:
    assert set(z.files) == {"model/h0/weights", "model/h0/biases", "model/out/weights", "model/out/biases",
                            "target_model/h0/weights", "target_model/h0/biases", "target_model/out/weights",
                            "target_model/out/biases", "__slots__", "__opt_state__"}
    assert z["model/h0/weights"].shape == (4, 3)
    np.testing.assert_array_equal(z["model/h0/weights"].reshape(-1), want["params"][:12].numpy())
  for b in e.buffers.values():
    b.add_(1.0)
  s.force_save()                                                 # a second, different checkpoint
  index = open(os.path.join(d, "checkpoint")).read().splitlines()
  second = index[0].split('"')[1]
  assert second != first and [l.split('"')[1] for l in index[1:]] == [first, second]
  # a fresh process picks up the latest one
  e2 = _HostEngine(2)
  util.SaverUtil(e2, d, save_freq=3600)
  for k in want:
    np.testing.assert_array_equal(e2.buffers[k].numpy(), (want[k] + 1.0).numpy())
  # and an explicit path restores the older one
  s2 = util.SaverUtil(e2, d, save_freq=3600)
  s2.restore(os.path.join(d, first))
  for k in want:
    np.testing.assert_array_equal(e2.buffers[k].numpy(), want[k].numpy())


def test_saver_rejects_a_checkpoint_of_another_model(tmp_path):
  d = str(tmp_path)
  util.SaverUtil(_HostEngine(1), d, save_freq=3600)
  other = _HostEngine(3)
  other.nets["model"]._shapes[0] = ("h0/weights", (3, 4))
  with pytest.raises(ValueError):
    util.SaverUtil(other, d, save_freq=3600)
  other = _HostEngine(3)
  other.nets["model"]._shapes[0] = ("h9/weights", (4, 3))
  with pytest.raises(KeyError):
    util.SaverUtil(other, d, save_freq=3600)


def test_save_if_required_follows_the_clock(tmp_path):
  s = util.SaverUtil(_HostEngine(1), str(tmp_path), save_freq=3600)
  n0 = len(os.listdir(str(tmp_path)))
  s.save_if_required()
  assert len(os.listdir(str(tmp_path))) == n0
  s.next_scheduled_save_time = 0
  s.save_if_required()
  assert len(os.listdir(str(tmp_path))) == n0 + 1


# ---- GPU: resume == uninterrupted ------------------------------------------------------------------------------------------

def _batches(shape, B, n, seed):
  from cartpoleplusplus_b200.replay_memory import Batch
  rs = np.random.RandomState(seed)
  out = []
  for _ in range(n):
    s1 = (rs.randint(0, 256, (B,) + shape).astype(np.float16) / np.float16(255)).astype(np.float16)
    s2 = (rs.randint(0, 256, (B,) + shape).astype(np.float16) / np.float16(255)).astype(np.float16)
    out.append(Batch(s1, rs.uniform(-1, 1, (B, 2)).astype(np.float32), np.ones((B, 1), np.float32),
                     (rs.rand(B, 1) > 0.1).astype(np.float32), s2))
  return out


@pytest.mark.gpu
def test_naf_adam_resume_is_bit_identical(tmp_path):
  from tests import gpu_util as U
  shape, B = (16, 16, 3, 1, 2), 16
  batches = _batches(shape, B, 6, 5)
  adam = {"learning_rate": 1e-3}

  def fresh():
    naf, nets, eng, _ = U.make_naf(shape, True, batch_size=B, optimiser="Adam", optimiser_args=adam)
    return naf, nets, eng

  naf, nets, eng = fresh()
  nets["target_value"].set_as_target_network_for(nets["value"], 0.1)
  for b in batches[:3]:
    naf.train(b)
    nets["target_value"].update_weights()
  saver = util.SaverUtil(eng, str(tmp_path), save_freq=3600)          # saves the state after 3 steps
  losses_a = []
  for b in batches[3:]:
    losses_a.append(naf.train(b))
    nets["target_value"].update_weights()
  end = {k: eng.buffers[k].clone() for k in ("params", "target_params", "slots", "opt_state")}

  naf2, nets2, eng2 = fresh()
  for part in eng2.nets:                                              # a different initialisation, to be overwritten
    eng2.part_view(part).normal_()
  nets2["target_value"].update_weights_op = nets2["target_value"]._create_variables_copy_op(nets2["value"], 0.1)
  util.SaverUtil(eng2, str(tmp_path), save_freq=3600)                 # finds and loads it
  losses_b = []
  for b in batches[3:]:
    losses_b.append(naf2.train(b))
    nets2["target_value"].update_weights()
  assert losses_a == losses_b
  for k in end:
    assert torch.equal(end[k], eng2.buffers[k]), k


@pytest.mark.gpu
def test_ddpg_restore_in_place_under_graph_replay(tmp_path):
  """the fused step replays a captured CUDA graph that holds the parameter pointers: restore must write in place"""
  from tests import gpu_util as U
  shape, B = (16, 16, 3, 1, 2), 16
  batches = _batches(shape, B, 5, 9)
  nets, eng, _ = U.make_ddpg(shape, True, batch_size=B)
  nets["target_actor"].set_as_target_network_for(nets["actor"], 0.1)
  nets["target_critic"].set_as_target_network_for(nets["critic"], 0.1)
  ptrs = {k: eng.buffers[k].data_ptr() for k in ("params", "target_params")}
  eng.train_step(batches[0])
  saver = util.SaverUtil(eng, str(tmp_path), save_freq=3600)
  a = []
  for b in batches[1:]:
    a.append(eng.train_step(b))
  end = eng.buffers["params"].clone()
  saver.restore(os.path.join(str(tmp_path), saver._latest()))
  assert ptrs == {k: eng.buffers[k].data_ptr() for k in ptrs}
  b2 = [eng.train_step(b) for b in batches[1:]]
  assert a == b2 and torch.equal(end, eng.buffers["params"])


def test_only_rank_zero_writes(tmp_path, monkeypatch):
  monkeypatch.setenv("RANK", "1"); monkeypatch.setenv("WORLD_SIZE", "2")
  s = util.SaverUtil(_HostEngine(1), str(tmp_path), save_freq=3600)      # nothing to load, and a non-zero rank does not save
  assert os.listdir(str(tmp_path)) == []
  s.force_save()
  assert os.listdir(str(tmp_path)) == []
  monkeypatch.setenv("RANK", "0")
  s.force_save()
  assert "checkpoint" in os.listdir(str(tmp_path))


def test_a_stray_rank_variable_does_not_silence_checkpoints(tmp_path, monkeypatch):
  """single-process run with a leftover RANK (old torchrun / MPI / scheduler environment): it IS the writer"""
  monkeypatch.setenv("RANK", "3"); monkeypatch.delenv("WORLD_SIZE", raising=False)
  util.SaverUtil(_HostEngine(1), str(tmp_path), save_freq=3600)
  assert "checkpoint" in os.listdir(str(tmp_path))
