#!/usr/bin/env python
"""Drop-in for the reference's ddpg_cartpole.py network classes, agent loop and command line
(/root/reference/ddpg_cartpole.py): ActorNetwork :78-145, CriticNetwork :148-248,
DeepDeterministicPolicyGradientAgent :251-409, flags :18-56, main :412-443.

All arithmetic (forward, dQ/da, TD target, loss, backward, clip, SGD, target update) runs in
libcartpolepp (cpp_ddpg_*, include/cartpolepp.h)."""
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

np.set_printoptions(precision=5, threshold=10000, suppress=True, linewidth=10000)


def build_parser():
  parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
  parser.add_argument('--num-eval', type=int, default=0, help="if >0 just run this many episodes with no training")
  parser.add_argument('--max-num-actions', type=int, default=0,
                      help="train for (at least) this number of actions (always finish current episode) ignore if <=0")
  parser.add_argument('--max-run-time', type=int, default=0,
                      help="train for (at least) this number of seconds (always finish current episode) ignore if <=0")
  parser.add_argument('--ckpt-dir', type=str, default=None, help="if set save ckpts to this dir")
  parser.add_argument('--ckpt-freq', type=int, default=3600, help="freq (sec) to save ckpts")
  parser.add_argument('--batch-size', type=int, default=128, help="training batch size")
  parser.add_argument('--batches-per-step', type=int, default=5, help="number of batches to train per step")
  parser.add_argument('--dont-do-rollouts', action="store_true",
                      help="by dft we do rollouts to generate data then train after each rollout. if this flag is set we"
                           " dont do any rollouts. this only makes sense to do if --event-log-in set.")
  parser.add_argument('--target-update-rate', type=float, default=0.0001,
                      help="affine combo for updating target networks each time we run a training batch")
  parser.add_argument('--use-batch-norm', action='store_true', help="whether to use batch norm on conv layers")
  parser.add_argument('--actor-hidden-layers', type=str, default="100,100,50", help="actor hidden layer sizes")
  parser.add_argument('--critic-hidden-layers', type=str, default="100,100,50", help="critic hidden layer sizes")
  parser.add_argument('--actor-learning-rate', type=float, default=0.001, help="learning rate for actor")
  parser.add_argument('--critic-learning-rate', type=float, default=0.01, help="learning rate for critic")
  parser.add_argument('--discount', type=float, default=0.99, help="discount for RHS of critic bellman equation update")
  parser.add_argument('--event-log-in', type=str, default=None, help="prepopulate replay memory with entries from this event log")
  parser.add_argument('--replay-memory-size', type=int, default=22000, help="max size of replay memory")
  parser.add_argument('--replay-memory-burn-in', type=int, default=1000, help="dont train from replay memory until it reaches this size")
  parser.add_argument('--eval-action-noise', action='store_true', help="whether to use noise during eval")
  parser.add_argument('--action-noise-theta', type=float, default=0.01,
                      help="OrnsteinUhlenbeckNoise theta (rate of change) param for action exploration")
  parser.add_argument('--action-noise-sigma', type=float, default=0.05,
                      help="OrnsteinUhlenbeckNoise sigma (magnitude) param for action exploration")
  util.add_opts(parser)
  from . import synthetic_env
  synthetic_env.add_opts(parser)
  return parser


# the reference keeps `opts` as a module global read by the network constructors (ddpg_cartpole.py:57,91)
opts = None
VERBOSE_DEBUG = False


def set_opts(o):
  global opts
  opts = o
  return o


def default_opts(argv=()):
  return build_parser().parse_args(list(argv))


class ActorNetwork(base_network.Network):
  """ the actor represents the learnt policy mapping states to actions"""

  def __init__(self, namespace, input_state, action_dim):
    super(ActorNetwork, self).__init__(namespace)
    self.input_state = input_state
    self.action_dim = action_dim
    self.exploration_noise = util.OrnsteinUhlenbeckNoise(action_dim, opts.action_noise_theta, opts.action_noise_sigma)
    opts.hidden_layers = opts.actor_hidden_layers                 # ddpg_cartpole.py:91
    final_hidden = self.input_state_network(self.input_state, opts)
    # action dim output. note: actors out is (-1, 1) and scaled in env as required.
    self.output_action = base_network.fully_connected(final_hidden, action_dim, scope='output_action', activation="tanh")
    self._finalise(self.output_action)
    self._critic = None

  def initial_flat(self, rng):
    return self.initial_values(rng, small_uniform=("output_action",))   # random_uniform(-0.001, 0.001) :94

  def init_ops_for_training(self, critic):
    # the gradient is d(output_action)/d(theta) weighted by -dQ/da from the critic, clipped, SGD (:111-119)
    self._critic = critic

  def action_given(self, state, add_noise=False):
    if self._part != "actor":      # the engine entry point reads the online parameters; the reference loops never ask the target actor
      raise NotImplementedError("action_given on %s: only the online actor is evaluated on this path" % self.namespace)
    actions = self._need_engine().action_given(np.asarray(state)[None])
    # NOTE: noise is added _outside_ the device graph, as in the reference (:127-134)
    if add_noise:
      actions[0] += self.exploration_noise.sample()
      actions = np.clip(1, -1, actions)  # reference quirk kept: rotated arguments (Appendix C-6)
    return actions

  def train(self, state):
    self._need_engine().actor_train(state)


class CriticNetwork(base_network.Network):
  """ the critic represents a mapping from state & actors action to a quality score."""

  def __init__(self, namespace, actor):
    super(CriticNetwork, self).__init__(namespace)
    self.actor = actor
    self.input_state = actor.input_state
    self.input_action = actor.output_action        # stop_gradient(actor.output_action) :162
    if opts.use_raw_pixels:
      conv_net = self.simple_conv_net_on(self.input_state, opts)
      # Appendix C-2: the trunk is flattened before hidden1 (what input_state_network does for the actor)
      hidden1 = base_network.fully_connected(base_network.flatten(conv_net), 200, scope='hidden1')
      hidden2 = base_network.fully_connected(hidden1, 50, scope='hidden2')
      concat_inputs = base_network.concat_action(hidden2, actor.action_dim)
      final_hidden = base_network.fully_connected(concat_inputs, 50, scope="hidden3")
    else:
      flat_input_state = base_network.flatten(self.input_state)
      concat_inputs = base_network.concat_action(flat_input_state, actor.action_dim)
      final_hidden = self.hidden_layers_starting_at(concat_inputs, opts.critic_hidden_layers)
    self.q_value = base_network.fully_connected(final_hidden, 1, scope='q_value', activation=None)
    self._finalise(self.q_value)
    self._target_critic = None

  def initial_flat(self, rng):
    return self.initial_values(rng)

  def init_ops_for_training(self, target_critic):
    self._target_critic = target_critic
    self.input_state_2 = target_critic.input_state
    DDPGEngine(self.actor, self, target_critic.actor, target_critic, opts)

  def q_gradients_wrt_actions(self):
    """ gradients for the q.value w.r.t just input_action; used for actor training"""
    return ("dq_da", self)

  def train(self, batch):
    self._need_engine().critic_train(batch)

  def check_loss(self, batch):
    return self._need_engine().check_loss(batch)


class DDPGEngine(EngineBase):
  """owns [actor|critic], [target_actor|target_critic] and the flat gradient buffer of one DDPG agent"""

  def __init__(self, actor, critic, target_actor, target_critic, o, seed=None):
    EngineBase.__init__(self)
    self.nets = dict(actor=actor, critic=critic, target_actor=target_actor, target_critic=target_critic)
    self.o = o
    self.max_batch = 0
    self.handle = None
    self._comm_init, self._p2p_prepare, self._p2p_connect = self.lib.cpp_ddpg_comm_init, self.lib.cpp_ddpg_p2p_prepare, self.lib.cpp_ddpg_p2p_connect
    self._layout()
    rng = np.random.RandomState(seed)
    for part, net in self.nets.items():
      net._engine, net._part = self, part
      self.part_view(part).copy_(torch.from_numpy(net.initial_flat(rng)))
    self._ensure(max(1, int(getattr(o, "batch_size", 128))))

  def _config(self, max_batch):
    cfg = _lib.DDPGConfig()
    cfg.actor, cfg.critic = self.nets["actor"]._spec, self.nets["critic"]._spec
    cfg.actor_lr, cfg.critic_lr = self.o.actor_learning_rate, self.o.critic_learning_rate
    cfg.discount = self.o.discount
    cfg.gradient_clip = self.o.gradient_clip if self.o.gradient_clip is not None else 0.0
    cfg.target_update_rate = self.o.target_update_rate
    cfg.max_batch, cfg.world_size, cfg.rank = max_batch, self.world_size, self.rank
    return cfg

  def _layout(self):
    h = C.c_void_p()
    _lib.check(self.lib.cpp_ddpg_create(C.byref(self._config(1)), C.byref(h)))
    out = (C.c_int64 * 5)()
    _lib.check(self.lib.cpp_ddpg_layout(h, out))
    self.lib.cpp_ddpg_destroy(h)
    self.n_actor, self.n_critic, self.off_critic, self.off_loss, self.total = [int(v) for v in out]
    dev = self.device
    self.buffers["params"] = torch.zeros(self.off_loss, dtype=torch.float32, device=dev)
    self.buffers["target_params"] = torch.zeros(self.off_loss, dtype=torch.float32, device=dev)
    self.buffers["grads"] = torch.zeros(self.total, dtype=torch.float32, device=dev)
    self.parts = dict(actor=("params", 0, self.n_actor), critic=("params", self.off_critic, self.n_critic),
                      target_actor=("target_params", 0, self.n_actor),
                      target_critic=("target_params", self.off_critic, self.n_critic))

  def _ensure(self, B):
    if B <= self.max_batch:
      return
    if self.handle is not None:
      torch.cuda.current_stream().synchronize()
      self.lib.cpp_ddpg_destroy(self.handle)
    h = C.c_void_p()
    _lib.check(self.lib.cpp_ddpg_create(C.byref(self._config(B)), C.byref(h)))
    nbytes = int(self.lib.cpp_ddpg_workspace_bytes(h))
    self.buffers["workspace"] = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
    b = _lib.DDPGBuffers()
    b.params, b.target_params, b.grads = (self.buffers[k].data_ptr() for k in ("params", "target_params", "grads"))
    b.workspace, b.workspace_bytes = self.buffers["workspace"].data_ptr(), nbytes
    _lib.check(self.lib.cpp_ddpg_bind(h, C.byref(b)))
    self.handle, self.max_batch = h, B
    if self.lib_comm:                                 # a re-created agent joins the communicator group again (all ranks grow together)
      self._connect()
    self.out_loss = torch.zeros(1, dtype=torch.float32, device=self.device)
    self.out_td = torch.zeros(B, dtype=torch.float32, device=self.device)
    self.out_q = torch.zeros(B, dtype=torch.float32, device=self.device)
    self.out_action = torch.zeros(B * self.nets["actor"].action_dim, dtype=torch.float32, device=self.device)

  def _batch_args(self, batch):
    s1, a, r, m, s2 = self._staged(batch)
    B = int(s1.shape[0])
    if state_flag(s1) != state_flag(s2):
      raise TypeError("state_1 and state_2 must share a dtype")
    return s1, a, r, m, s2, B

  def actor_backward(self, state):
    """forward + dQ/da + actor backward -> grads[actor part] (unclipped batch-sum gradient)"""
    s1 = self.stage("s1", state)
    B = int(s1.shape[0])
    self._ensure(B)
    _lib.check(self.lib.cpp_ddpg_actor_backward(self.handle, _lib.ptr(s1), state_flag(s1), B, B * self.world_size, self._stream()))

  def actor_apply(self):
    _lib.check(self.lib.cpp_ddpg_actor_apply(self.handle, self._stream()))

  def actor_train(self, state):
    self.actor_backward(state)
    if self.dp is not None:
      self.dp.all_reduce_sum(self.buffers["grads"][:self.off_critic])
    self.actor_apply()

  def critic_backward(self, batch, reuse_s1_trunk=False):
    """target forward + critic forward + TD/MSE + critic backward -> grads[critic part], grads[off_loss]"""
    s1, a, r, m, s2, B = self._batch_args(batch)
    self._ensure(B)
    _lib.check(self.lib.cpp_ddpg_critic_backward(self.handle, _lib.ptr(s1), _lib.ptr(a), _lib.ptr(r), _lib.ptr(m), _lib.ptr(s2),
                                                 state_flag(s1), B, B * self.world_size, 1 if reuse_s1_trunk else 0, self._stream()))

  def critic_apply(self):
    _lib.check(self.lib.cpp_ddpg_critic_apply(self.handle, self._stream()))

  def critic_train(self, batch, reuse_s1_trunk=False):
    self.critic_backward(batch, reuse_s1_trunk)
    if self.dp is not None:
      self.dp.all_reduce_sum(self.buffers["grads"][self.off_critic:])
    self.critic_apply()

  def train_step(self, batch, moments=None):
    """one DDPG grad-step = actor.train(batch.state_1); critic.train(batch) (ddpg_cartpole.py:332-334) with the
    batch staged once and the critic trunk on state_1 computed once.  moments: optional (mean_inv_s1, mean_inv_s2)
    device tensors with the whitening statistics of the GLOBAL batch (data parallel / replay-resident path)."""
    s1, a, r, m, s2, B = self._batch_args(batch)
    self._ensure(B)
    self._need_global_moments(moments, self.nets["actor"]._spec.pixels)
    st = self._stream()
    if moments is not None:
      _lib.check(self.lib.cpp_ddpg_set_moments(self.handle, _lib.ptr(moments[0]), _lib.ptr(moments[1])))
    Bg = B * self.world_size
    if (self.dp is None and self.world_size == 1) or self.lib_comm:
      # one call, one graph: backward of both networks, (data parallel: the in-library all-reduce,) clip + SGD of both
      _lib.check(self.lib.cpp_ddpg_train_step(self.handle, _lib.ptr(s1), _lib.ptr(a), _lib.ptr(r), _lib.ptr(m), _lib.ptr(s2),
                                              state_flag(s1), B, st))
      if moments is not None:
        _lib.check(self.lib.cpp_ddpg_set_moments(self.handle, None, None))
      self._release_slot()
      return
    _lib.check(self.lib.cpp_ddpg_step_backward(self.handle, _lib.ptr(s1), _lib.ptr(a), _lib.ptr(r), _lib.ptr(m), _lib.ptr(s2),
                                               state_flag(s1), B, Bg, st))
    if self.dp is not None:
      self.dp.all_reduce_sum(self.buffers["grads"])        # the single gradient all-reduce of the step
    _lib.check(self.lib.cpp_ddpg_step_apply(self.handle, st))
    self._release_slot()
    if moments is not None:
      _lib.check(self.lib.cpp_ddpg_set_moments(self.handle, None, None))

  def update_targets(self):
    """target_actor.update_weights(); target_critic.update_weights() (ddpg_cartpole.py:336-337) in one launch"""
    _lib.check(self.lib.cpp_ddpg_update_targets(self.handle, C.c_float(self.o.target_update_rate), self._stream()))

  def last_loss(self):
    """device->host read of the loss the last critic step wrote (grads[off_loss])"""
    return float(self.buffers["grads"][self.off_loss].item())

  def check_loss(self, batch):
    s1, a, r, m, s2, B = self._batch_args(batch)
    self._ensure(B)
    _lib.check(self.lib.cpp_ddpg_check_loss(self.handle, _lib.ptr(s1), _lib.ptr(a), _lib.ptr(r), _lib.ptr(m), _lib.ptr(s2),
                                            state_flag(s1), B, _lib.ptr(self.out_loss), _lib.ptr(self.out_td),
                                            _lib.ptr(self.out_q), self._stream()))
    return (float(self.out_loss.item()), self.out_td[:B].cpu().numpy().reshape(B, 1),
            self.out_q[:B].cpu().numpy().reshape(B, 1))

  def action_given(self, states):
    return self._action_given(self.lib.cpp_ddpg_action_given_fast, self.lib.cpp_ddpg_action_given, states, self.nets["actor"].action_dim)


class DeepDeterministicPolicyGradientAgent(object):
  def __init__(self, env):
    self.env = env
    state_shape = self.env.observation_space.shape
    action_dim = self.env.action_space.shape[1]
    self.replay_memory = replay_memory.ReplayMemory(opts.replay_memory_size, state_shape, action_dim)
    s1 = base_network.Placeholder(state_shape, "s1")
    s2 = base_network.Placeholder(state_shape, "s2")
    self.actor = ActorNetwork("actor", s1, action_dim)
    self.critic = CriticNetwork("critic", self.actor)
    self.target_actor = ActorNetwork("target_actor", s2, action_dim)
    self.target_critic = CriticNetwork("target_critic", self.target_actor)
    self.actor.init_ops_for_training(self.critic)
    self.critic.init_ops_for_training(self.target_critic)

  def post_var_init_setup(self):
    if opts.event_log_in:
      self.replay_memory.reset_from_event_log(opts.event_log_in)
    self.target_actor.set_as_target_network_for(self.actor, opts.target_update_rate)
    self.target_critic.set_as_target_network_for(self.critic, opts.target_update_rate)

  def run_training(self, max_num_actions, max_run_time, batch_size, batches_per_step, saver_util=None):
    start_time = time.time()
    num_actions_taken = 0
    n = 0
    while True:
      rewards = []
      losses = []
      if opts.dont_do_rollouts:
        pass
      else:
        state_1 = self.env.reset()
        initial_state = np.copy(state_1)
        action_reward_state_sequence = []
        done = False
        while not done:
          action = self.actor.action_given(state_1, add_noise=True)
          state_2, reward, done, _ = self.env.step(action)
          rewards.append(reward)
          action_reward_state_sequence.append((action, reward, np.copy(state_2)))
          state_1 = state_2
        self.replay_memory.add_episode(initial_state, action_reward_state_sequence)

      if self.replay_memory.size() > opts.replay_memory_burn_in:
        for _ in range(batches_per_step):
          batch = self.replay_memory.batch(batch_size)
          # == self.actor.train(batch.state_1); self.critic.train(batch), staged once
          self.actor._engine.train_step(batch)
        self.target_actor.update_weights()
        self.target_critic.update_weights()
        if VERBOSE_DEBUG:
          td_loss, td, q_value = self.critic.check_loss(batch)
          print("temporal_difference_loss", td_loss)
          print("temporal_difference", td.T)
          print("q_value", q_value.T)

      stats = collections.OrderedDict()
      stats["time"] = time.time()
      stats["n"] = n
      stats["mean_losses"] = float(np.mean(losses)) if losses else float("nan")   # never filled (Appendix C-4)
      stats["total_reward"] = float(np.sum(rewards))
      stats["episode_len"] = len(rewards)
      stats["replay_memory_stats"] = self.replay_memory.current_stats()
      self.actor._engine.check_piece_overflow()      # (one device sync per STATS line)
      print("STATS %s\t%s" % (datetime.datetime.now().strftime('%Y-%m-%d %H:%M:%S'), json.dumps(stats)))
      sys.stdout.flush()
      n += 1

      if saver_util is not None:
        saver_util.save_if_required()
      if VERBOSE_DEBUG or n % 10 == 0:
        self.run_eval(1)

      num_actions_taken += len(rewards)
      if max_num_actions > 0 and num_actions_taken > max_num_actions:
        break
      if max_run_time > 0 and time.time() > start_time + max_run_time:
        break
      if opts.dont_do_rollouts and max_num_actions > 0 and n * batches_per_step * batch_size > max_num_actions:
        break   # pure-training runs take no actions: bound them by trained transitions instead

  def run_eval(self, num_episodes, add_noise=False):
    """ run num_episodes of eval and output episode length and rewards """
    for i in range(num_episodes):
      state = self.env.reset()
      total_reward = 0
      steps = 0
      done = False
      while not done:
        action = self.actor.action_given(state, add_noise)
        state, reward, done, _ = self.env.step(action)
        print("EVALSTEP r%s %s %s %s %s" % (i, steps, np.squeeze(action), np.linalg.norm(action), reward))
        total_reward += reward
        steps += 1
      print("EVAL", i, steps, total_reward)
    sys.stdout.flush()


def main(argv=None):
  from . import synthetic_env
  set_opts(build_parser().parse_args(argv))
  sys.stderr.write("%s\n" % opts)
  env = synthetic_env.SyntheticCartpole(opts=opts, discrete_actions=False)
  agent = DeepDeterministicPolicyGradientAgent(env=env)
  for net in (agent.actor, agent.critic, agent.target_actor, agent.target_critic):
    for v in net._variables():
      sys.stderr.write("%s %s\n" % (v.name, util.shape_and_product_of(v.shape)))
  # setup saver util and either load latest ckpt or keep the initialised variables (ddpg_cartpole.py:419-422)
  saver_util = None
  if opts.ckpt_dir is not None:
    saver_util = util.SaverUtil(agent.actor._engine, opts.ckpt_dir, opts.ckpt_freq)
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
