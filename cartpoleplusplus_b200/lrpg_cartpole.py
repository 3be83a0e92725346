#!/usr/bin/env python
"""Drop-in for the reference's lrpg_cartpole.py (/root/reference/lrpg_cartpole.py): the likelihood-ratio
policy-gradient agent on low-dim state (graph :80-130, rollout :136-163, train :165-182, loop :187-253).
BASELINE config 1 "CPU plumbing": the same FC / loss / clip / optimiser kernels at tiny sizes (cpp_lrpg_*).
Action sampling (tf.multinomial, :95-96) stays on the host."""
import argparse
import collections
import ctypes as C
import datetime
import json
import sys
import time
import numpy as np
import torch

from . import _lib, base_network, util
from ._engine import EngineBase


def build_parser():
  parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
  parser.add_argument('--num-eval', type=int, default=0, help="if >0 just run this many episodes with no training")
  parser.add_argument('--max-num-actions', type=int, default=0)
  parser.add_argument('--max-run-time', type=int, default=0)
  parser.add_argument('--ckpt-dir', type=str, default=None)
  parser.add_argument('--ckpt-freq', type=int, default=3600)
  parser.add_argument('--hidden-layers', type=str, default="100,50", help="hidden layer sizes")
  parser.add_argument('--learning-rate', type=float, default=0.0001, help="unused, as in the reference (Appendix C-3)")
  parser.add_argument('--num-train-batches', type=int, default=10, help="number of training batches to run")
  parser.add_argument('--rollouts-per-batch', type=int, default=10, help="number of rollouts to run for each training batch")
  parser.add_argument('--eval-action-noise', action='store_true')
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


class LikelihoodRatioPolicyGradientAgent(base_network.Network):

  def __init__(self, env):
    base_network.Network.__init__(self, "model")     # (the reference forgets this call, Appendix C-3)
    assert not opts.use_raw_pixels, "TODO: add convnet from ddpg here"      # lrpg_cartpole.py:42
    self.env = env
    num_actions = self.env.action_space.n
    self.num_actions = num_actions
    self.observations = base_network.Placeholder(self.env.observation_space.shape, "observations")
    flat_input_state = base_network.flatten(self.observations)
    final_hidden = self.hidden_layers_starting_at(flat_input_state, opts.hidden_layers)
    logits = base_network.fully_connected(final_hidden, num_actions, scope="fully_connected", activation=None)
    self._finalise(logits)
    self.optimiser = util.construct_optimiser(opts)
    LRPGEngine(self, opts)

  def initial_flat(self, rng):
    return self.initial_values(rng)

  def logits_given(self, observations):
    return self._engine.logits(observations)

  def sample_action_given(self, observation, doing_eval=False):
    """ sample one action given observation"""
    if not doing_eval and np.random.random() < 0.1:        # epsilon greedy, lrpg_cartpole.py:141-142
      return self.env.action_space.sample()
    z = self.logits_given(np.asarray(observation, dtype=np.float32)[None])[0].astype(np.float64)
    p = np.exp(z - z.max()); p /= p.sum()
    return int(np.random.choice(self.num_actions, p=p))     # tf.multinomial(logits, 1), host side

  def rollout(self, doing_eval=False):
    observations, actions, rewards = [], [], []
    observation = self.env.reset()
    done = False
    while not done:
      observations.append(observation)
      action = self.sample_action_given(observation, doing_eval)
      observation, reward, done, _ = self.env.step(action)
      actions.append(action)
      rewards.append(reward)
    return observations, actions, rewards

  def train(self, observations, actions, advantages):
    """ take one training step given observations, actions and subsequent advantages"""
    return self._engine.train(observations, actions, advantages)

  def post_var_init_setup(self):
    pass

  def run_training(self, max_num_actions, max_run_time, rollouts_per_batch, saver_util=None):
    start_time = time.time()
    num_actions_taken = 0
    n = 0
    while True:
      total_rewards = []
      batch_observations, batch_actions, batch_advantages = [], [], []
      for _ in range(rollouts_per_batch):
        observations, actions, rewards = self.rollout()
        batch_observations += observations
        batch_actions += actions
        batch_advantages += [sum(rewards)] * len(rewards)       # lrpg_cartpole.py:209
        total_rewards.append(sum(rewards))
      losses = []
      if min(total_rewards) == max(total_rewards):               # :213-216: standardising equal advantages would divide by zero
        sys.stderr.write("converged? standardisation of advantaged will barf here....\n")
        loss = 0
      else:
        loss = self.train(batch_observations, batch_actions, batch_advantages)
        losses.append(loss)
      # the reference's STATS keys and accounting (:221-250): mean over the (zero or one) losses of this iteration - NaN
      # when training was skipped, exactly like np.mean([]) there -, the episode length of the LAST rollout, and
      # num_actions_taken advanced by that last rollout only
      stats = collections.OrderedDict()
      stats["time"] = time.time()
      stats["n"] = n
      stats["mean_losses"] = float(np.mean(losses)) if losses else float("nan")
      stats["total_reward"] = float(np.sum(total_rewards))
      stats["episode_len"] = len(rewards)
      print("STATS %s\t%s" % (datetime.datetime.now().strftime('%Y-%m-%d %H:%M:%S'), json.dumps(stats)))
      sys.stdout.flush()
      n += 1
      if saver_util is not None:
        saver_util.save_if_required()
      if VERBOSE_DEBUG or n % 10 == 0:                           # :239-240 occasional eval
        self.run_eval(1)
      num_actions_taken += len(rewards)
      if max_num_actions > 0 and num_actions_taken > max_num_actions:
        break
      if max_run_time > 0 and time.time() > start_time + max_run_time:
        break

  def run_eval(self, num_eval):
    for _ in range(num_eval):
      _, _, rewards = self.rollout(doing_eval=True)
      print(sum(rewards))


class LRPGEngine(EngineBase):
  def __init__(self, agent, o, seed=None):
    EngineBase.__init__(self)
    self.agent, self.o = agent, o
    self.kind, self.hp = agent.optimiser
    self.n = agent.num_params()
    npad = (self.n + 3) // 4 * 4
    dev = self.device
    self.buffers["params"] = torch.zeros(npad, dtype=torch.float32, device=dev)
    self.buffers["grads"] = torch.zeros(npad, dtype=torch.float32, device=dev)
    self.buffers["slots"] = torch.zeros(max(4, self.kind * npad), dtype=torch.float32, device=dev)
    self.buffers["opt_state"] = torch.ones(4, dtype=torch.float32, device=dev)
    self.parts = dict(model=("params", 0, self.n))
    agent._engine, agent._part = self, "model"
    self.part_view("model").copy_(torch.from_numpy(agent.initial_flat(np.random.RandomState(seed))))
    self.max_batch, self.handle = 0, None
    self._ensure(4096)

  def _ensure(self, N):
    if N <= self.max_batch:
      return
    if self.handle is not None:
      torch.cuda.current_stream().synchronize()
      self.lib.cpp_lrpg_destroy(self.handle)
    cfg = _lib.LRPGConfig()
    cfg.model = self.agent._spec
    cfg.gradient_clip = self.o.gradient_clip if self.o.gradient_clip is not None else 0.0
    cfg.optimiser = self.kind
    cfg.lr, cfg.momentum, cfg.beta1, cfg.beta2, cfg.eps = (self.hp[k] for k in ("lr", "momentum", "beta1", "beta2", "eps"))
    cfg.max_batch = N
    h = C.c_void_p()
    _lib.check(self.lib.cpp_lrpg_create(C.byref(cfg), C.byref(h)))
    nbytes = int(self.lib.cpp_lrpg_workspace_bytes(h))
    self.buffers["workspace"] = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
    b = _lib.LRPGBuffers()
    b.params, b.grads, b.slots, b.opt_state = (self.buffers[k].data_ptr() for k in ("params", "grads", "slots", "opt_state"))
    b.workspace, b.workspace_bytes = self.buffers["workspace"].data_ptr(), nbytes
    _lib.check(self.lib.cpp_lrpg_bind(h, C.byref(b)))
    self.handle, self.max_batch = h, N
    self.out_logits = torch.zeros(N * self.agent.num_actions, dtype=torch.float32, device=self.device)

  def train(self, observations, actions, advantages):
    obs = self.stage("obs", np.asarray(observations, dtype=np.float32), torch.float32)
    act = self.stage("act", np.asarray(actions, dtype=np.int32), torch.int32)
    adv = self.stage("adv", np.asarray(advantages, dtype=np.float32), torch.float32)
    N = int(obs.shape[0])
    self._ensure(N)
    loss = C.c_float()
    _lib.check(self.lib.cpp_lrpg_train(self.handle, _lib.ptr(obs), _lib.ptr(act), _lib.ptr(adv), N, C.byref(loss), self._stream()))
    return float(loss.value)

  def logits(self, observations):
    obs = self.stage("obs", np.asarray(observations, dtype=np.float32), torch.float32)
    N = int(obs.shape[0])
    self._ensure(N)
    K = self.agent.num_actions
    _lib.check(self.lib.cpp_lrpg_logits(self.handle, _lib.ptr(obs), N, _lib.ptr(self.out_logits), self._stream()))
    return self.out_logits[:N * K].cpu().numpy().reshape(N, K)


def main(argv=None):
  from . import synthetic_env
  set_opts(build_parser().parse_args(argv))
  sys.stderr.write("%s\n" % opts)
  env = synthetic_env.SyntheticCartpole(opts=opts, discrete_actions=True)
  agent = LikelihoodRatioPolicyGradientAgent(env=env)
  saver_util = None
  if opts.ckpt_dir is not None:     # lrpg_cartpole.py:278-281
    saver_util = util.SaverUtil(agent._engine, opts.ckpt_dir, opts.ckpt_freq)
  agent.post_var_init_setup()
  if opts.num_eval > 0:
    agent.run_eval(opts.num_eval)
  else:
    agent.run_training(opts.max_num_actions, opts.max_run_time, opts.rollouts_per_batch, saver_util)
    if saver_util is not None:
      saver_util.force_save()


if __name__ == "__main__":
  main()
