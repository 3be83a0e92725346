#!/usr/bin/env python
"""Drop-in for the reference's naf_cartpole.py networks, agent loop and command line
(/root/reference/naf_cartpole.py): ValueNetwork :93-114, NafNetwork :117-284,
NormalizedAdvantageFunctionAgent :287-455, flags :18-71.  Arithmetic: libcartpolepp cpp_naf_*."""
import argparse
import collections
import ctypes as C
import datetime
import json
import sys
import time
import numpy as np
import torch

from . import _lib, base_network, replay_memory, util
from ._engine import EngineBase, state_flag


def build_parser():
  parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
  parser.add_argument('--num-eval', type=int, default=0, help="if >0 just run this many episodes with no training")
  parser.add_argument('--max-num-actions', type=int, default=0)
  parser.add_argument('--max-run-time', type=int, default=0)
  parser.add_argument('--ckpt-dir', type=str, default=None)
  parser.add_argument('--ckpt-freq', type=int, default=3600)
  parser.add_argument('--batch-size', type=int, default=128, help="training batch size")
  parser.add_argument('--batches-per-step', type=int, default=5, help="number of batches to train per step")
  parser.add_argument('--dont-do-rollouts', action="store_true")
  parser.add_argument('--target-update-rate', type=float, default=0.0001)
  parser.add_argument('--use-batch-norm', action='store_true')
  parser.add_argument('--share-input-state-representation', action='store_true',
                      help="whether to share the input state network between V, mu and L")
  parser.add_argument('--hidden-layers', type=str, default="100,50", help="hidden layer sizes")
  parser.add_argument('--discount', type=float, default=0.99)
  parser.add_argument('--event-log-in', type=str, default=None)
  parser.add_argument('--replay-memory-size', type=int, default=22000)
  parser.add_argument('--replay-memory-burn-in', type=int, default=1000)
  parser.add_argument('--eval-action-noise', action='store_true')
  parser.add_argument('--action-noise-theta', type=float, default=0.01)
  parser.add_argument('--action-noise-sigma', type=float, default=0.05)
  parser.add_argument('--gpu-mem-fraction', type=float, default=None, help="accepted and ignored (TF option)")
  util.add_opts(parser)
  from . import synthetic_env
  synthetic_env.add_opts(parser)
  return parser


opts = None
VERBOSE_DEBUG = False


def set_opts(o):
  global opts
  opts = o
  return o


def default_opts(argv=()):
  return build_parser().parse_args(list(argv))


class ValueNetwork(base_network.Network):
  """ Value network component of a NAF network. Created as seperate net because it has a target network."""

  def __init__(self, namespace, input_state, hidden_layer_config):
    super(ValueNetwork, self).__init__(namespace)
    self.input_state = input_state
    self.input_state_representation = self.input_state_network(input_state, opts)
    self.value = base_network.fully_connected(self.input_state_representation, 1, scope='fc', activation=None)
    self._finalise(self.value)

  def initial_flat(self, rng):
    return self.initial_values(rng)

  def value_given(self, state):
    if self._part != "value":      # the engine entry point reads the online parameters; the reference loops never ask a target net
      raise NotImplementedError("value_given on %s: only the online value network is evaluated on this path" % self.namespace)
    return self._need_engine().value_given(state)


class _SubNet(base_network.Network):
  """naf/output_action and naf/l_values sub-networks (variable scopes inside NafNetwork)"""

  def __init__(self, namespace, input_state, num_outputs, activation, small):
    super(_SubNet, self).__init__(namespace)
    rep = self.input_state_network(input_state, opts)
    self._finalise(base_network.fully_connected(rep, num_outputs, scope='fc', activation=activation))
    self._small = small

  def initial_flat(self, rng):
    return self.initial_values(rng, small_uniform=("fc",) if self._small else ())


class _HeadNet(base_network.Network):
  """--share-input-state-representation (naf_cartpole.py:151-154,176-179): the sub-network is only its `fc` layer, on top of
  value_net.input_state_representation; the gradient it sends into the representation reaches the value/* variables"""

  def __init__(self, namespace, representation, num_outputs, activation, small):
    super(_HeadNet, self).__init__(namespace)
    rep_dim = representation.fc[-1][1] if representation.fc else int(np.prod(representation.feature_shape()))
    src = base_network.Placeholder((rep_dim,), "input_state_representation")
    self._finalise(base_network.fully_connected(src, num_outputs, scope='fc', activation=activation))
    self._small = small

  def initial_flat(self, rng):
    return self.initial_values(rng, small_uniform=("fc",) if self._small else ())


class NafNetwork(base_network.Network):

  def __init__(self, namespace, input_state, input_state_2, value_net, target_value_net, action_dim):
    super(NafNetwork, self).__init__(namespace)
    self.exploration_noise = util.OrnsteinUhlenbeckNoise(action_dim, opts.action_noise_theta, opts.action_noise_sigma)
    self.value_net, self.target_value_net = value_net, target_value_net
    self.input_state, self.input_state_2 = input_state, input_state_2
    self.action_dim = action_dim
    # mu (output_action): its own input_state_network + tanh head with U(+-1e-3) weights (:150-161)
    num_l_values = (action_dim * (action_dim + 1)) // 2
    self.share = bool(getattr(opts, "share_input_state_representation", False))
    if self.share:
      rep = value_net.input_state_representation
      self.mu_net = _HeadNet(namespace + "/output_action", rep, action_dim, "tanh", True)
      self.l_net = _HeadNet(namespace + "/l_values", rep, num_l_values, None, False)
    else:
      self.mu_net = _SubNet(namespace + "/output_action", input_state, action_dim, "tanh", True)
      # l_values: lower-triangular entries, diagonal exponentiated in the head kernel (:172-207)
      self.l_net = _SubNet(namespace + "/l_values", input_state, num_l_values, None, False)
    self.optimiser = util.construct_optimiser(opts)
    NAFEngine(self, opts)

  def _variables(self):
    return self.mu_net._variables() + self.l_net._variables()

  def action_given(self, state, add_noise):
    actions = self._need_engine().action_given(np.asarray(state)[None])
    if add_noise:
      actions[0] += self.exploration_noise.sample()
      actions = np.clip(1, -1, actions)   # reference quirk kept (Appendix C-6)
    return actions

  def train(self, batch):
    return self._need_engine().train(batch)

  def debug_values(self, batch):
    return self._need_engine().debug_values(batch)


class NAFEngine(EngineBase):
  def __init__(self, naf, o, seed=None):
    EngineBase.__init__(self)
    self.naf, self.o = naf, o
    self.nets = collections.OrderedDict([("value", naf.value_net), ("mu", naf.mu_net), ("l", naf.l_net),
                                         ("target_value", naf.target_value_net)])
    self.kind, self.hp = naf.optimiser
    self.max_batch, self.handle = 0, None
    self._comm_init, self._p2p_prepare, self._p2p_connect = self.lib.cpp_naf_comm_init, self.lib.cpp_naf_p2p_prepare, self.lib.cpp_naf_p2p_connect
    self._layout()
    rng = np.random.RandomState(seed)
    for part, net in self.nets.items():
      net._engine, net._part = self, part
      self.part_view(part).copy_(torch.from_numpy(net.initial_flat(rng)))
    naf._engine = self
    self._ensure(max(1, int(getattr(o, "batch_size", 128))))

  def _config(self, max_batch):
    cfg = _lib.NAFConfig()
    cfg.value, cfg.mu, cfg.l = self.nets["value"]._spec, self.nets["mu"]._spec, self.nets["l"]._spec
    cfg.discount = self.o.discount
    cfg.gradient_clip = self.o.gradient_clip if self.o.gradient_clip is not None else 0.0
    cfg.target_update_rate = self.o.target_update_rate
    cfg.optimiser = self.kind
    cfg.lr, cfg.momentum, cfg.beta1, cfg.beta2, cfg.eps = (self.hp[k] for k in ("lr", "momentum", "beta1", "beta2", "eps"))
    cfg.max_batch, cfg.action_dim = max_batch, self.naf.action_dim
    cfg.world_size, cfg.rank = self.world_size, self.rank
    cfg.share_input_state_representation = 1 if self.naf.share else 0
    return cfg

  def _layout(self):
    h = C.c_void_p()
    _lib.check(self.lib.cpp_naf_create(C.byref(self._config(1)), C.byref(h)))
    out = (C.c_int64 * 7)()
    _lib.check(self.lib.cpp_naf_layout(h, out))
    self.lib.cpp_naf_destroy(h)
    self.n_v, self.n_m, self.n_l, self.off_m, self.off_l, self.off_loss, self.total = [int(v) for v in out]
    dev = self.device
    self.buffers["params"] = torch.zeros(self.off_loss, dtype=torch.float32, device=dev)
    self.buffers["target_params"] = torch.zeros(self.off_m, dtype=torch.float32, device=dev)
    self.buffers["grads"] = torch.zeros(self.total, dtype=torch.float32, device=dev)
    self.buffers["slots"] = torch.zeros(max(4, self.kind * self.off_loss), dtype=torch.float32, device=dev)
    self.buffers["opt_state"] = torch.ones(4, dtype=torch.float32, device=dev)      # beta1^0, beta2^0
    self.parts = dict(value=("params", 0, self.n_v), mu=("params", self.off_m, self.n_m), l=("params", self.off_l, self.n_l),
                      target_value=("target_params", 0, self.n_v))

  def _ensure(self, B):
    if B <= self.max_batch:
      return
    if self.handle is not None:
      torch.cuda.current_stream().synchronize()
      self.lib.cpp_naf_destroy(self.handle)
    h = C.c_void_p()
    _lib.check(self.lib.cpp_naf_create(C.byref(self._config(B)), C.byref(h)))
    nbytes = int(self.lib.cpp_naf_workspace_bytes(h))
    self.buffers["workspace"] = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
    b = _lib.NAFBuffers()
    b.params, b.target_params, b.grads, b.slots, b.opt_state = (
        self.buffers[k].data_ptr() for k in ("params", "target_params", "grads", "slots", "opt_state"))
    b.workspace, b.workspace_bytes = self.buffers["workspace"].data_ptr(), nbytes
    _lib.check(self.lib.cpp_naf_bind(h, C.byref(b)))
    self.handle, self.max_batch = h, B
    if self.lib_comm:
      self._connect()
    A = self.naf.action_dim
    dev = self.device
    self.out = dict(l=torch.zeros(B * (A * (A + 1)) // 2, dtype=torch.float32, device=dev), loss=torch.zeros(1, dtype=torch.float32, device=dev),
                    V=torch.zeros(B, dtype=torch.float32, device=dev), A=torch.zeros(B, dtype=torch.float32, device=dev),
                    V2=torch.zeros(B, dtype=torch.float32, device=dev), act=torch.zeros(B * A, dtype=torch.float32, device=dev))

  def _batch_args(self, batch):
    s1, a, r, m, s2 = self._staged(batch)
    return s1, a, r, m, s2, int(s1.shape[0])

  def backward(self, batch):
    """forward + head + backward of value/mu/l -> grads (unclipped), grads[off_loss] = loss"""
    s1, a, r, m, s2, B = self._batch_args(batch)
    self._ensure(B)
    _lib.check(self.lib.cpp_naf_backward(self.handle, _lib.ptr(s1), _lib.ptr(a), _lib.ptr(r), _lib.ptr(m), _lib.ptr(s2),
                                         state_flag(s1), B, B * self.world_size, self._stream()))

  def apply(self, check=True):
    loss = C.c_float()
    _lib.check(self.lib.cpp_naf_apply(self.handle, 1 if check else 0, C.byref(loss), self._stream()))
    return float(loss.value)

  def apply_async(self):
    """clip + optimiser without the host read-back of the loss: the update is skipped on the device when l_values / L / loss
    were non-finite and the error surfaces at the next last_loss() (steady-state loops, bench.py)"""
    _lib.check(self.lib.cpp_naf_apply(self.handle, 2, None, self._stream()))
    return None

  def last_loss(self):
    lf = self.buffers["grads"][self.off_loss:self.off_loss + 2].cpu().numpy()
    if lf[1] != 0 or not np.isfinite(lf[0]):
      raise _lib.CppError(-4, "check_numerics: non-finite l_values / L / loss (naf_cartpole.py:242-245)")
    return float(lf[0])

  def update_targets(self):
    """target_value_net.update_weights() (naf_cartpole.py:373) as one launch"""
    _lib.check(self.lib.cpp_naf_update_targets(self.handle, C.c_float(self.o.target_update_rate), self._stream()))

  def train(self, batch, moments=None, sync=True):
    """naf.train(batch) (naf_cartpole.py:264-272).  moments: optional (mean_inv_s1, mean_inv_s2) device tensors with the
    whitening statistics of the GLOBAL batch (data parallel: every rank trains on a slice of it)"""
    self._need_global_moments(moments, self.nets["value"]._spec.pixels)
    if moments is not None:
      _lib.check(self.lib.cpp_naf_set_moments(self.handle, _lib.ptr(moments[0]), _lib.ptr(moments[1])))
    self.backward(batch)                                   # data parallel with lib_comm: the all-reduce ran inside
    if moments is not None:
      _lib.check(self.lib.cpp_naf_set_moments(self.handle, None, None))
    if self.dp is not None and not self.lib_comm:
      self.dp.all_reduce_sum(self.buffers["grads"])        # the single gradient all-reduce of the step
    loss = self.apply(True) if sync else self.apply_async()
    self._release_slot()
    return loss

  def debug_values(self, batch):
    s1, a, r, m, s2, B = self._batch_args(batch)
    self._ensure(B)
    o = self.out
    _lib.check(self.lib.cpp_naf_debug_values(self.handle, _lib.ptr(s1), _lib.ptr(a), _lib.ptr(r), _lib.ptr(m), _lib.ptr(s2),
                                             state_flag(s1), B, _lib.ptr(o["l"]), _lib.ptr(o["loss"]), _lib.ptr(o["V"]),
                                             _lib.ptr(o["A"]), _lib.ptr(o["V2"]), self._stream()))
    A = self.naf.action_dim
    NL = (A * (A + 1)) // 2
    vals = [o["l"][:B * NL].cpu().numpy().reshape(B, NL), o["loss"].cpu().numpy().reshape(()), o["V"][:B].cpu().numpy().reshape(B, 1),
            o["A"][:B].cpu().numpy().reshape(B, 1), o["V2"][:B].cpu().numpy().reshape(B, 1)]
    return [np.squeeze(v) for v in vals]

  def action_given(self, states):
    return self._action_given(self.lib.cpp_naf_action_given_fast, self.lib.cpp_naf_action_given, states, self.naf.action_dim)

  def value_given(self, states):
    s = self.stage("s_act", states)
    B = int(s.shape[0]); self._ensure(B)
    _lib.check(self.lib.cpp_naf_value_given(self.handle, _lib.ptr(s), state_flag(s), B, _lib.ptr(self.out["V"]), self._stream()))
    return self.out["V"][:B].cpu().numpy().reshape(B, 1)


class NormalizedAdvantageFunctionAgent(object):
  def __init__(self, env):
    self.env = env
    state_shape = self.env.observation_space.shape
    action_dim = self.env.action_space.shape[1]
    self.replay_memory = replay_memory.ReplayMemory(opts.replay_memory_size, state_shape, action_dim)
    s1 = base_network.Placeholder(state_shape, "s1")
    s2 = base_network.Placeholder(state_shape, "s2")
    self.value_net = ValueNetwork("value", s1, opts.hidden_layers)
    self.target_value_net = ValueNetwork("target_value", s2, opts.hidden_layers)
    self.naf = NafNetwork("naf", s1, s2, self.value_net, self.target_value_net, action_dim)

  def post_var_init_setup(self):
    if opts.event_log_in:
      self.replay_memory.reset_from_event_log(opts.event_log_in)
    self.target_value_net.set_as_target_network_for(self.value_net, opts.target_update_rate)

  def run_training(self, max_num_actions, max_run_time, batch_size, batches_per_step, saver_util=None):
    start_time = time.time()
    num_actions_taken = 0
    n = 0
    while True:
      rewards, losses = [], []
      if not opts.dont_do_rollouts:
        state_1 = self.env.reset()
        initial_state = np.copy(state_1)
        action_reward_state_sequence = []
        done = False
        while not done:
          action = self.naf.action_given(state_1, add_noise=True)
          state_2, reward, done, _ = self.env.step(action)
          rewards.append(reward)
          action_reward_state_sequence.append((action, reward, np.copy(state_2)))
          state_1 = state_2
        self.replay_memory.add_episode(initial_state, action_reward_state_sequence)
      if self.replay_memory.size() > opts.replay_memory_burn_in:
        for _ in range(batches_per_step):
          batch = self.replay_memory.batch(batch_size)
          losses.append(self.naf.train(batch))
        self.target_value_net.update_weights()
      stats = collections.OrderedDict()
      stats["time"] = time.time()
      stats["n"] = n
      stats["mean_losses"] = float(np.mean(losses)) if losses else float("nan")
      stats["total_reward"] = float(np.sum(rewards))
      stats["episode_len"] = len(rewards)
      stats["replay_memory_stats"] = self.replay_memory.current_stats()
      self.naf._engine.check_piece_overflow()      # (one device sync per STATS line)
      print("STATS %s\t%s" % (datetime.datetime.now().strftime('%Y-%m-%d %H:%M:%S'), json.dumps(stats)))
      sys.stdout.flush()
      if saver_util is not None:
        saver_util.save_if_required()
      n += 1
      if VERBOSE_DEBUG or n % 10 == 0:
        self.run_eval(1)
      num_actions_taken += len(rewards)
      if max_num_actions > 0 and num_actions_taken > max_num_actions:
        break
      if max_run_time > 0 and time.time() > start_time + max_run_time:
        break
      if opts.dont_do_rollouts and max_num_actions > 0 and n * batches_per_step * batch_size > max_num_actions:
        break

  def run_eval(self, num_episodes, add_noise=False):
    for i in range(num_episodes):
      state = self.env.reset()
      total_reward, steps, done = 0, 0, False
      while not done:
        action = self.naf.action_given(state, add_noise)
        state, reward, done, _ = self.env.step(action)
        print("EVALSTEP e%d s%d action=%s (l2=%s) => reward %s" % (i, steps, action, np.linalg.norm(action), reward))
        total_reward += reward
        steps += 1
      print("EVAL", i, steps, total_reward)
    sys.stdout.flush()


def main(argv=None):
  from . import synthetic_env
  set_opts(build_parser().parse_args(argv))
  sys.stderr.write("%s\n" % opts)
  env = synthetic_env.SyntheticCartpole(opts=opts, discrete_actions=False)
  agent = NormalizedAdvantageFunctionAgent(env=env)
  # setup saver util and either load latest ckpt or keep the initialised variables (naf_cartpole.py:467-470)
  saver_util = None
  if opts.ckpt_dir is not None:
    saver_util = util.SaverUtil(agent.naf._engine, opts.ckpt_dir, opts.ckpt_freq)
  agent.post_var_init_setup()
  if opts.num_eval > 0:
    agent.run_eval(opts.num_eval, opts.eval_action_noise)
  else:
    agent.run_training(opts.max_num_actions, opts.max_run_time, opts.batch_size, opts.batches_per_step, saver_util)
    if saver_util is not None:
      saver_util.force_save()
  env.reset()


if __name__ == "__main__":
  main()
