"""Oracle restatement of the reference's TensorFlow-0.x/slim training graphs in torch (CPU).

PARITY UNPINNED: the arithmetic lives in ``tensorflow`` + ``tensorflow.contrib.slim``
(imported at /root/reference/base_network.py:4-5; no version pin exists in the reference, API
vintage => TF r0.10-r0.11).  Neither is in /root/reference nor in this image and the reference
holds no numeric test for this path, so this file restates the *published* op semantics
 and is anchored on the reference's call sites cited per function.
fp64 = truth; fp32 = the "reference CPU path" that bench.py times.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Optional graph variants restated here: --share-input-state-representation (naf_cartpole.py:151-154,176-179),
--use-dropout (base_network.py:69-70; masks are inputs, see `dropout`) and --use-batch-norm (base_network.py:74-79 + slim.batch_norm defaults, Appendix A-5: batch statistics under the global
IS_TRAINING flag in every network of a train op, never-updated moving statistics in inference).

Reference defects at HEAD are resolved as SURVEY.md Appendix C fixes them (C-1: opts=None means
no dropout; C-2: the pixel critic flattens the conv trunk before hidden1).
"""
import collections
import math
import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------- definitions

FC = collections.namedtuple("FC", "scope out act")          # act in {'relu','tanh',None}


class NetDef(object):
  """One reference network: optional conv trunk -> flatten -> FC stack (optional action concat)."""

  def __init__(self, ns, state_shape, pixels, fc, concat_at=None, action_dim=0, batch_norm=False):
    self.ns, self.state_shape, self.pixels = ns, tuple(state_shape), bool(pixels)
    self.batch_norm = bool(batch_norm) and bool(pixels)      # --use-batch-norm only touches the conv layers, base_network.py:74-79
    self.fc, self.concat_at, self.action_dim = list(fc), concat_at, action_dim
    if pixels:
      H, W = state_shape[0], state_shape[1]
      self.H, self.W = H, W
      self.cin = int(np.prod(state_shape[2:]))      # (H,W,3,C,R) -> 3*C*R, base_network.py:88-90
      h, w = H, W
      for _ in range(3):
        h, w = h // 2, w // 2                       # max_pool2d 2x2/2 VALID, base_network.py:107,115,123
      self.feat = h * w * 10
    else:
      self.cin = 0
      self.feat = int(np.prod(state_shape))         # slim.flatten, base_network.py:133

  def var_shapes(self):
    """[(name, shape)] in TF variable-creation order."""
    out = []
    if self.pixels:
      cin = self.cin
      for name, k in (("conv1", 5), ("conv2", 5), ("conv3", 3)):   # base_network.py:103-123
        out.append(("%s/%s/weights" % (self.ns, name), (k, k, cin, 10)))
        if self.batch_norm:
          # slim.conv2d with a normalizer_fn creates no bias; slim.batch_norm (center=True, scale=False) creates beta and the
          # two moving statistics, in this order (Appendix A-4, A-5)
          out.append(("%s/%s/BatchNorm/beta" % (self.ns, name), (10,)))
          out.append(("%s/%s/BatchNorm/moving_mean" % (self.ns, name), (10,)))
          out.append(("%s/%s/BatchNorm/moving_variance" % (self.ns, name), (10,)))
        else:
          out.append(("%s/%s/biases" % (self.ns, name), (10,)))
        cin = 10
    d = self.feat
    for i, l in enumerate(self.fc):
      if self.concat_at == i:
        d += self.action_dim
      out.append(("%s/%s/weights" % (self.ns, l.scope), (d, l.out)))
      out.append(("%s/%s/biases" % (self.ns, l.scope), (l.out,)))
      d = l.out
    return out

  def num_params(self):
    return sum(int(np.prod(s)) for _, s in self.var_shapes())

  def trainable_names(self):
    """tf.trainable_variables of the namespace: the moving statistics are variables (the target update copies them,
    base_network.py:25-26) but not trainable"""
    return [n for n, _ in self.var_shapes() if "/moving_" not in n]


def _hidden(sizes):
  if isinstance(sizes, str):
    sizes = [int(s) for s in sizes.split(",")]          # base_network.py:60-61
  return [FC("h%d" % i, s, "relu") for i, s in enumerate(sizes)]   # base_network.py:63-68


def ddpg_actor(ns, state_shape, pixels, hidden="100,100,50", action_dim=2, batch_norm=False):
  """ddpg_cartpole.py:90-100"""
  return NetDef(ns, state_shape, pixels, _hidden(hidden) + [FC("output_action", action_dim, "tanh")], batch_norm=batch_norm)


def ddpg_critic(ns, state_shape, pixels, hidden="100,100,50", action_dim=2, batch_norm=False):
  """ddpg_cartpole.py:161-184 (pixel branch repaired per Appendix C-2)"""
  if pixels:
    fc = [FC("hidden1", 200, "relu"), FC("hidden2", 50, "relu"), FC("hidden3", 50, "relu"),
          FC("q_value", 1, None)]
    return NetDef(ns, state_shape, True, fc, concat_at=2, action_dim=action_dim, batch_norm=batch_norm)
  return NetDef(ns, state_shape, False, _hidden(hidden) + [FC("q_value", 1, None)],
                concat_at=0, action_dim=action_dim)


def naf_value(ns, state_shape, pixels, hidden="100,50", batch_norm=False):
  """naf_cartpole.py:96-109"""
  return NetDef(ns, state_shape, pixels, _hidden(hidden) + [FC("fc", 1, None)], batch_norm=batch_norm)


def naf_mu(state_shape, pixels, hidden="100,50", action_dim=2, batch_norm=False):
  """naf_cartpole.py:147-161"""
  return NetDef("naf/output_action", state_shape, pixels, _hidden(hidden) + [FC("fc", action_dim, "tanh")], batch_norm=batch_norm)


def naf_l(state_shape, pixels, hidden="100,50", action_dim=2, batch_norm=False):
  """naf_cartpole.py:172-184"""
  return NetDef("naf/l_values", state_shape, pixels,
                _hidden(hidden) + [FC("fc", action_dim * (action_dim + 1) // 2, None)], batch_norm=batch_norm)


def naf_shared_heads(rep_dim, action_dim=2):
  """--share-input-state-representation, naf_cartpole.py:151-154,176-179: output_action / l_values are only their `fc` layer,
  fed by value_net.input_state_representation (width rep_dim)"""
  return (NetDef("naf/output_action", (rep_dim,), False, [FC("fc", action_dim, "tanh")]),
          NetDef("naf/l_values", (rep_dim,), False, [FC("fc", action_dim * (action_dim + 1) // 2, None)]))


def lrpg_model(state_shape, hidden="100,50", num_actions=5):
  """lrpg_cartpole.py:80-88"""
  return NetDef("model", state_shape, False, _hidden(hidden) + [FC("fully_connected", num_actions, None)])


def init_params(netdef, rs, dtype=torch.float64):
  """xavier-uniform weights / zero biases (slim defaults, Appendix A-4/A-6); action heads
  U(+-1e-3) (ddpg_cartpole.py:94, naf_cartpole.py:155).  The reference never seeds TF, so
  parity tests always inject weights; this is only a generator of plausible values."""
  P = collections.OrderedDict()
  for name, shape in netdef.var_shapes():
    if name.endswith("biases") or name.endswith("/beta") or name.endswith("/moving_mean"):
      v = np.zeros(shape)                       # zeros_initializer (slim defaults)
    elif name.endswith("/moving_variance"):
      v = np.ones(shape)                        # ones_initializer
    elif len(shape) == 4:
      kh, kw, ci, co = shape
      lim = math.sqrt(6.0 / (kh * kw * ci + kh * kw * co))
      v = rs.uniform(-lim, lim, shape)
    else:
      lim = math.sqrt(6.0 / (shape[0] + shape[1]))
      if name.endswith("output_action/weights") or name == "naf/output_action/fc/weights":
        lim = 1e-3
      v = rs.uniform(-lim, lim, shape)
    P[name] = torch.tensor(v.astype(np.float32), dtype=dtype)   # values are fp32-representable
  return P


def retarget(P, src_ns, dst_ns):
  """same tensors under the target namespace (base_network.py:28)"""
  return collections.OrderedDict((dst_ns + k[len(src_ns):], v.clone()) for k, v in P.items())


def flat(P, names=None):
  names = names or list(P.keys())
  return torch.cat([P[n].reshape(-1) for n in names])


# ----------------------------------------------------------------------------- forward

def whiten(x):
  """base_network.py:95-99: per-channel batch moments (population variance), eps 1e-6,
  x*inv - mean*inv (tf.nn.batch_normalization with scale=offset=None)."""
  mean = x.mean(dim=(0, 1, 2))
  var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
  inv = torch.rsqrt(var + 1e-6)
  return x * inv - mean * inv


# base_network.py:11: ONE global placeholder fed to every Session.run - True in the train ops, False in action_given /
# check_loss / debug_values.  It reaches slim.batch_norm of EVERY network evaluated by that run, the target networks too.
IS_TRAINING = True


class is_training(object):
  """with is_training(False): ... - the feed_dict={IS_TRAINING: ...} of one Session.run"""

  def __init__(self, value):
    self.value = bool(value)

  def __enter__(self):
    global IS_TRAINING
    self.saved, IS_TRAINING = IS_TRAINING, self.value

  def __exit__(self, *exc):
    global IS_TRAINING
    IS_TRAINING = self.saved


BN_EPSILON = 1e-3      # slim.batch_norm default epsilon; decay 0.999 is irrelevant: UPDATE_OPS never run (Appendix A-5)


def batch_norm(x, beta, moving_mean, moving_variance):
  """slim.batch_norm(center=True, scale=False, epsilon=1e-3) on an NCHW tensor (Appendix A-5).  Training: per-channel batch
  mean and POPULATION variance over (B, H, W), gradients flow through both.  Inference: the moving statistics, which the
  reference never updates (it never runs the UPDATE_OPS collection), i.e. their initial 0 / 1 or whatever a target copy put there."""
  if IS_TRAINING:
    mean = x.mean(dim=(0, 2, 3))
    var = ((x - mean[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))
  else:
    mean, var = moving_mean, moving_variance
  inv = torch.rsqrt(var + BN_EPSILON)
  return (x - mean[None, :, None, None]) * inv[None, :, None, None] + beta[None, :, None, None]


# Routing pinned from outside (parity tests only): relu + max_pool2d is a piecewise-linear map whose pieces are selected by
# discrete decisions (which of the four window positions wins, whether the ReLU is open).  A decision that sits within
# rounding of a tie may legitimately be taken differently by an fp32 / tensor-core forward pass and by this fp64 graph; the
# gradient behind it then differs by a whole term, not by rounding.  `with gates({(ns, "conv1"): u8 [B][PH][PW][10], ...})`
# evaluates relu(max_pool(.)) of those layers with the GIVEN decisions (0..3 = dy*2+dx of the winner, 4 = ReLU closed) and
# records in GATE_STATS how many decisions differ from the ones this graph would have taken and how far each of those is
# from a tie - so that a test can assert "same function, and every disagreement is a tie within rounding".
GATES = None
GATE_STATS = None


class gates(object):
  """with gates(routing) as stats: ... - stats[(ns, layer)] = dict(n, mismatched, worst_gap_rel)"""

  def __init__(self, routing):
    self.routing = routing

  def __enter__(self):
    global GATES, GATE_STATS
    self.saved = (GATES, GATE_STATS)
    GATES, GATE_STATS = self.routing, {}
    return GATE_STATS

  def __exit__(self, *exc):
    global GATES, GATE_STATS
    GATES, GATE_STATS = self.saved


def _relu_pool_pinned(x, amax, key):
  """x (B, C, H, W) pre-activation, amax u8 (B, H//2, W//2, C) -> (B, C, H//2, W//2) with the routing of `amax`"""
  B, C, H, W = x.shape
  ph, pw = H // 2, W // 2
  win = x[:, :, :2 * ph, :2 * pw].reshape(B, C, ph, 2, pw, 2).permute(0, 1, 2, 4, 3, 5).reshape(B, C, ph, pw, 4)
  a = torch.as_tensor(np.asarray(amax)).to(torch.int64).reshape(B, ph, pw, C).permute(0, 3, 1, 2)
  is_open = (a < 4).to(x.dtype)
  picked = torch.gather(win, 4, a.clamp(max=3).unsqueeze(-1)).squeeze(-1)
  out = picked * is_open
  with torch.no_grad():
    best, arg = win.max(dim=4)
    own = torch.where(best > 0, arg, torch.full_like(arg, 4))
    diff = own != a
    # distance of a disagreeing decision from the tie: |value the pinned routing passes on - value this graph passes on|
    gap = (out - F.relu(best)).abs()
    scale = float(x.abs().max())
    st = GATE_STATS.setdefault(key, dict(n=0, mismatched=0, worst_gap_rel=0.0))
    st["n"] += int(diff.numel()); st["mismatched"] += int(diff.sum())
    st["worst_gap_rel"] = max(st["worst_gap_rel"], float(gap.max()) / max(scale, 1e-30))
  return out


def conv_trunk(nd, P, state):
  """base_network.py:73-127 -> (B, h, w, 10) NHWC"""
  B = state.shape[0]
  x = state.reshape(B, nd.H, nd.W, nd.cin)             # plain reshape, Appendix A-2
  x = whiten(x).permute(0, 3, 1, 2)
  for name, k in (("conv1", 5), ("conv2", 5), ("conv3", 3)):
    W = P["%s/%s/weights" % (nd.ns, name)].permute(3, 2, 0, 1)   # HWIO -> OIHW
    if nd.batch_norm:                                             # no bias; BN in front of the ReLU (Appendix A-4)
      pre = "%s/%s/BatchNorm/" % (nd.ns, name)
      x = F.conv2d(x, W, None, stride=1, padding=k // 2)
      x = batch_norm(x, P[pre + "beta"], P[pre + "moving_mean"], P[pre + "moving_variance"])
    else:
      b = P["%s/%s/biases" % (nd.ns, name)]
      x = F.conv2d(x, W, b, stride=1, padding=k // 2)             # SAME, cross-correlation
    if GATES is not None and (nd.ns, name) in GATES:
      x = _relu_pool_pinned(x, GATES[(nd.ns, name)], (nd.ns, name))
    else:
      x = F.max_pool2d(F.relu(x), 2)                              # relu, then stride 2, VALID (floor)
  return x.permute(0, 2, 3, 1)


# --use-dropout (base_network.py:69-70): slim.dropout(keep_prob=0.5, is_training=IS_TRAINING) after every layer that
# hidden_layers_starting_at creates (scopes h0, h1, ...; NOT the pixel critic's hidden1-3, ddpg_cartpole.py:168-171, and not the
# heads).  TF's random stream cannot be reproduced, so the masks are INPUTS here: {(namespace, scope): 0/1 tensor [B][out]},
# one per graph node and Session.run - a node that feeds several consumers (the shared NAF representation) has ONE mask.
DROPOUT_KEEP_PROB = 0.5
DROPOUT_MASKS = None


class dropout(object):
  """with dropout(masks): ... - the masks the dropout ops of one Session.run draw (None: --use-dropout is off)"""

  def __init__(self, masks):
    self.masks = masks

  def __enter__(self):
    global DROPOUT_MASKS
    self.saved, DROPOUT_MASKS = DROPOUT_MASKS, self.masks

  def __exit__(self, *exc):
    global DROPOUT_MASKS
    DROPOUT_MASKS = self.saved


def has_dropout(scope):
  return len(scope) >= 2 and scope[0] == "h" and scope[1:].isdigit()


def draw_dropout_masks(rs, nets, B, dtype=torch.float64):
  """Bernoulli(keep_prob) masks for every dropout node of `nets` (a stand-in for TF's stream in tests and golden vectors)"""
  out = {}
  for nd in nets:
    for l in nd.fc:
      if has_dropout(l.scope):
        out[(nd.ns, l.scope)] = torch.tensor((rs.rand(B, l.out) < DROPOUT_KEEP_PROB).astype(np.float64), dtype=dtype)
  return out


def forward(nd, P, state, action=None, dtype=None, return_hidden=False, end_fc=None):
  dtype = dtype or next(iter(P.values())).dtype
  state = torch.as_tensor(np.asarray(state)) if not torch.is_tensor(state) else state
  x = state.to(dtype)                                  # fp16 replay states are cast by the feed
  B = x.shape[0]
  if nd.pixels:
    x = conv_trunk(nd, P, x).reshape(B, -1)            # flatten row-major (h,w,c), Appendix A-1
  else:
    x = x.reshape(B, -1)
  for i, l in enumerate(nd.fc):
    if end_fc is not None and i >= end_fc:            # stop in front of FC layer end_fc (the shared representation)
      break
    if nd.concat_at == i:
      x = torch.cat([x, action.to(dtype)], dim=1)     # ddpg_cartpole.py:170,175
    x = x @ P["%s/%s/weights" % (nd.ns, l.scope)] + P["%s/%s/biases" % (nd.ns, l.scope)]
    if l.act == "relu":
      x = F.relu(x)
    elif l.act == "tanh":
      x = torch.tanh(x)
    if DROPOUT_MASKS is not None and IS_TRAINING and has_dropout(l.scope):
      x = x * DROPOUT_MASKS[(nd.ns, l.scope)].to(dtype) / DROPOUT_KEEP_PROB      # inverted dropout: kept units scaled by 1/keep_prob
  return x


def _names(nd):
  return [n for n, _ in nd.var_shapes()]


def _trainable(nd):
  return nd.trainable_names()


def _leaf(P, names):
  for n in names:
    P[n] = P[n].detach().clone().requires_grad_(True)
  return [P[n] for n in names]


# ----------------------------------------------------------------------------- clip / optimisers

def global_norm(grads):
  return torch.sqrt(sum((g * g).sum() for g in grads))


def clip_by_global_norm(grads, clip):
  """util.py:45-50 -> tf.clip_by_global_norm: g * clip*min(1/norm, 1/clip)  (Appendix A-10)"""
  if clip is None:
    return list(grads), None
  norm = global_norm(grads)
  scale = clip * torch.minimum(1.0 / norm, torch.tensor(1.0 / clip, dtype=norm.dtype))
  return [g * scale for g in grads], norm


class Optimiser(object):
  """util.py:73-76 -> tf.train.{GradientDescent,Momentum,Adam}Optimizer (Appendix A-9)"""

  def __init__(self, kind, learning_rate, momentum=0.0, beta1=0.9, beta2=0.999, epsilon=1e-8):
    self.kind, self.lr = kind, learning_rate
    self.momentum, self.b1, self.b2, self.eps = momentum, beta1, beta2, epsilon
    self.t, self.slots = 0, {}

  def apply(self, params, grads):
    """in-place on the list of tensors `params`"""
    self.t += 1
    for i, (p, g) in enumerate(zip(params, grads)):
      if self.kind == "GradientDescent":
        p -= self.lr * g
      elif self.kind == "Momentum":
        acc = self.slots.setdefault(i, torch.zeros_like(p))
        acc.mul_(self.momentum).add_(g)
        p -= self.lr * acc
      elif self.kind == "Adam":
        m, v = self.slots.setdefault(i, (torch.zeros_like(p), torch.zeros_like(p)))
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        m += (1.0 - self.b1) * (g - m)
        v += (1.0 - self.b2) * (g * g - v)
        p -= lr_t * m / (torch.sqrt(v) + self.eps)        # epsilon-hat placement
      else:
        raise ValueError(self.kind)


def soft_update(target, source, coeff):
  """base_network.py:31: t <- t - c*(t - s), evaluated in that association"""
  return target - coeff * (target - source)


# ----------------------------------------------------------------------------- DDPG

def ddpg_actor_grads(actor, critic, P, s1):
  """ddpg_cartpole.py:102-119,220-222: grads = d mu/d theta . (-dQ(s1,mu(s1))/da), batch SUM."""
  names = _trainable(actor)
  leaves = _leaf(P, names)
  mu = forward(actor, P, s1)
  a_in = mu.detach().clone().requires_grad_(True)        # stop_gradient(actor.output_action) :162
  q = forward(critic, P, s1, a_in)
  dqda, = torch.autograd.grad(q.sum(), a_in)
  grads = torch.autograd.grad(mu, leaves, grad_outputs=-dqda)
  for n in names:
    P[n] = P[n].detach()
  return [g.detach() for g in grads], mu.detach(), q.detach(), dqda.detach()


def ddpg_critic_loss(critic, tactor, tcritic, P, batch, discount):
  """ddpg_cartpole.py:186-209 -> (loss, td, q)"""
  s1, a, r, mask, s2 = batch
  dt = next(iter(P.values())).dtype
  with torch.no_grad():
    mu2 = forward(tactor, P, s2)
    q2 = forward(tcritic, P, s2, mu2)
    y = torch.as_tensor(r).to(dt) + torch.as_tensor(mask).to(dt) * discount * q2
  q = forward(critic, P, s1, torch.as_tensor(a).to(dt))
  td = q - y
  return (td ** 2).mean(), td, q


def ddpg_critic_grads(critic, tactor, tcritic, P, batch, discount):
  names = _trainable(critic)
  leaves = _leaf(P, names)
  loss, td, q = ddpg_critic_loss(critic, tactor, tcritic, P, batch, discount)
  grads = torch.autograd.grad(loss, leaves)
  for n in names:
    P[n] = P[n].detach()
  return [g.detach() for g in grads], loss.detach(), td.detach(), q.detach()


class DDPGOracle(object):
  """The reference DDPG inner loop body (ddpg_cartpole.py:331-337) on explicit weights."""

  def __init__(self, state_shape, pixels, P, actor_hidden="100,100,50", critic_hidden="100,100,50",
               action_dim=2, actor_lr=1e-3, critic_lr=1e-2, discount=0.99, clip=5.0, tau=1e-4, batch_norm=False):
    self.actor = ddpg_actor("actor", state_shape, pixels, actor_hidden, action_dim, batch_norm)
    self.critic = ddpg_critic("critic", state_shape, pixels, critic_hidden, action_dim, batch_norm)
    self.tactor = ddpg_actor("target_actor", state_shape, pixels, actor_hidden, action_dim, batch_norm)
    self.tcritic = ddpg_critic("target_critic", state_shape, pixels, critic_hidden, action_dim, batch_norm)
    self.P = P
    self.actor_lr, self.critic_lr, self.discount, self.clip, self.tau = actor_lr, critic_lr, discount, clip, tau

  def actor_train(self, s1):
    g, mu, q, dqda = ddpg_actor_grads(self.actor, self.critic, self.P, s1)
    gc, norm = clip_by_global_norm(g, self.clip)
    for n, gi in zip(_trainable(self.actor), gc):
      self.P[n] = self.P[n] - self.actor_lr * gi
    return dict(grads=g, clipped=gc, norm=norm, mu=mu, q=q, dqda=dqda)

  def critic_train(self, batch):
    g, loss, td, q = ddpg_critic_grads(self.critic, self.tactor, self.tcritic, self.P, batch, self.discount)
    gc, norm = clip_by_global_norm(g, self.clip)
    for n, gi in zip(_trainable(self.critic), gc):
      self.P[n] = self.P[n] - self.critic_lr * gi
    return dict(grads=g, clipped=gc, norm=norm, loss=loss, td=td, q=q)

  def check_loss(self, batch):
    with torch.no_grad(), is_training(False):                                          # ddpg_cartpole.py:239-248
      return ddpg_critic_loss(self.critic, self.tactor, self.tcritic, self.P, batch, self.discount)

  def update_targets(self, coeff=None):
    c = self.tau if coeff is None else coeff
    for src, dst in ((self.actor, self.tactor), (self.critic, self.tcritic)):
      for ns_, nd_ in zip(_names(src), _names(dst)):
        self.P[nd_] = soft_update(self.P[nd_], self.P[ns_], c)

  def action_given(self, state):
    with torch.no_grad(), is_training(False):                                          # ddpg_cartpole.py:121-138
      return forward(self.actor, self.P, torch.as_tensor(np.asarray(state))[None])


# ----------------------------------------------------------------------------- NAF

def naf_quantities(value, mu_net, l_net, P, s1, action, action_dim, share=False):
  """naf_cartpole.py:147-221 -> (l_values, L, V, mu, A, Q)"""
  dt = next(iter(P.values())).dtype
  V = forward(value, P, s1)
  if share:      # the same graph nodes feed all three heads, so their gradients add up in value/* (:151-154,176-179)
    rep = forward(value, P, s1, end_fc=len(value.fc) - 1)
    mu = forward(mu_net, P, rep)
    l = forward(l_net, P, rep)
  else:
    mu = forward(mu_net, P, s1)
    l = forward(l_net, P, s1)
  B = l.shape[0]
  rows = []
  for r in range(action_dim):                               # naf_cartpole.py:195-203
    off = r * (r + 1) // 2
    lower = l[:, off:off + r]
    diag = torch.exp(l[:, off + r:off + r + 1])
    upper = torch.zeros(B, action_dim - r - 1, dtype=dt)
    rows.append(torch.cat([lower, diag, upper], dim=1))
  L = torch.stack(rows, dim=1)                              # (B, A, A)
  Pm = L @ L.transpose(1, 2)                                # :212
  d = (torch.as_tensor(action).to(dt) - mu).unsqueeze(-1)   # :166-167
  A = (-0.5 * (d.transpose(1, 2) @ (Pm @ d))).reshape(-1, 1)   # :215-218
  return l, L, V, mu, A, V + A


class NAFOracle(object):
  """naf_cartpole.py:367-373 body: naf.train(batch) then (every batches_per_step) target update."""

  def __init__(self, state_shape, pixels, P, hidden="100,50", action_dim=2, discount=0.99, clip=5.0,
               tau=1e-4, optimiser="GradientDescent", optimiser_args=None, share=False, batch_norm=False):
    self.value = naf_value("value", state_shape, pixels, hidden, batch_norm)
    self.tvalue = naf_value("target_value", state_shape, pixels, hidden, batch_norm)
    self.share = bool(share)
    if self.share:
      self.mu, self.l = naf_shared_heads(self.value.fc[-2].out if len(self.value.fc) > 1 else self.value.feat, action_dim)
    else:
      self.mu = naf_mu(state_shape, pixels, hidden, action_dim, batch_norm)
      self.l = naf_l(state_shape, pixels, hidden, action_dim, batch_norm)
    self.P, self.A, self.discount, self.clip, self.tau = P, action_dim, discount, clip, tau
    self.opt = Optimiser(optimiser, **(optimiser_args or {"learning_rate": 1e-3}))
    self.train_names = _trainable(self.value) + _trainable(self.mu) + _trainable(self.l)

  def _loss(self, batch):
    s1, a, r, mask, s2 = batch
    dt = next(iter(self.P.values())).dtype
    l, L, V, mu, A, Q = naf_quantities(self.value, self.mu, self.l, self.P, s1, a, self.A, self.share)
    with torch.no_grad():
      V2 = forward(self.tvalue, self.P, s2)
      y = torch.as_tensor(r).to(dt) + torch.as_tensor(mask).to(dt) * self.discount * V2   # :225-227
    loss = ((Q - y) ** 2).mean()                                                        # :230
    return loss, l, L, V, A, V2

  def train(self, batch):
    leaves = _leaf(self.P, self.train_names)
    loss, l, L, V, A, V2 = self._loss(batch)
    grads = [g.detach() for g in torch.autograd.grad(loss, leaves)]
    for n in self.train_names:
      self.P[n] = self.P[n].detach()
    if not (torch.isfinite(l).all() and torch.isfinite(L).all() and torch.isfinite(loss)):
      raise FloatingPointError("check_numerics")                                        # :242-245
    gc, norm = clip_by_global_norm(grads, self.clip)
    params = [self.P[n] for n in self.train_names]
    self.opt.apply(params, gc)
    return dict(loss=loss.detach(), grads=grads, clipped=gc, norm=norm)

  def debug_values(self, batch):
    with torch.no_grad(), is_training(False):                                          # naf_cartpole.py:274-284
      loss, l, L, V, A, V2 = self._loss(batch)
    return [np.squeeze(v.numpy()) for v in (l, loss, V, A, V2)]                         # :274-284

  def update_targets(self, coeff=None):
    c = self.tau if coeff is None else coeff
    for ns_, nd_ in zip(_names(self.value), _names(self.tvalue)):
      self.P[nd_] = soft_update(self.P[nd_], self.P[ns_], c)

  def action_given(self, state):
    with torch.no_grad(), is_training(False):                                          # naf_cartpole.py:247-262
      x = torch.as_tensor(np.asarray(state))[None]
      if self.share:
        x = forward(self.value, self.P, x, end_fc=len(self.value.fc) - 1)
      return forward(self.mu, self.P, x)


# ----------------------------------------------------------------------------- LRPG

def standardise(t):
  """util.py:37-43 (population std, no epsilon)"""
  mean = t.mean()
  return (t - mean) / torch.sqrt(((t - mean) ** 2).mean())


class LRPGOracle(object):
  """lrpg_cartpole.py:80-130,165-182"""

  def __init__(self, state_shape, P, hidden="100,50", num_actions=5, clip=5.0,
               optimiser="GradientDescent", optimiser_args=None):
    self.model = lrpg_model(state_shape, hidden, num_actions)
    self.P, self.clip, self.num_actions = P, clip, num_actions
    self.opt = Optimiser(optimiser, **(optimiser_args or {"learning_rate": 1e-3}))
    self.names = _names(self.model)

  def logits(self, observations):
    with torch.no_grad():
      return forward(self.model, self.P, observations)

  def train(self, observations, actions, advantages):
    dt = next(iter(self.P.values())).dtype
    leaves = _leaf(self.P, self.names)
    logits = forward(self.model, self.P, observations)
    logp = F.log_softmax(logits, dim=1)
    mask = F.one_hot(torch.as_tensor(np.asarray(actions), dtype=torch.int64), self.num_actions).to(dt)
    alp = (logp * mask).sum(dim=1)
    loss = -(alp * standardise(torch.as_tensor(np.asarray(advantages)).to(dt))).sum()
    grads = [g.detach() for g in torch.autograd.grad(loss, leaves)]
    for n in self.names:
      self.P[n] = self.P[n].detach()
    gc, norm = clip_by_global_norm(grads, self.clip)
    self.opt.apply([self.P[n] for n in self.names], gc)
    return dict(loss=loss.detach(), grads=grads, clipped=gc, norm=norm, logits=logits.detach())
