"""Independent checks of the torch oracle's semantics (parity is unpinned by the reference, so the
restatement is cross-checked against explicit numpy loops and closed forms from SURVEY.md App. A)."""
import json
import os
import numpy as np
import torch
import pytest

from oracle import nets_oracle as no


def naive_trunk(x, Ws, bs):
  """explicit loops: whitening, SAME cross-correlation, ReLU, 2x2/2 VALID max-pool"""
  x = x.astype(np.float64)
  mean = x.mean(axis=(0, 1, 2)); var = ((x - mean) ** 2).mean(axis=(0, 1, 2))
  x = (x - mean) / np.sqrt(var + 1e-6)
  for W, b in zip(Ws, bs):
    k = W.shape[0]; p = k // 2
    B, H, Wd, C = x.shape
    xp = np.zeros((B, H + 2 * p, Wd + 2 * p, C)); xp[:, p:p + H, p:p + Wd] = x
    y = np.zeros((B, H, Wd, W.shape[3]))
    for ky in range(k):
      for kx in range(k):
        y += np.einsum("bhwc,co->bhwo", xp[:, ky:ky + H, kx:kx + Wd], W[ky, kx])
    y = np.maximum(y + b, 0)
    h2, w2 = H // 2, Wd // 2
    y = y[:, :h2 * 2, :w2 * 2].reshape(B, h2, 2, w2, 2, -1).max(axis=(2, 4))
    x = y
  return x


@pytest.mark.parametrize("shape", [(16, 16, 3, 1, 2), (22, 18, 3, 2, 1)])
def test_trunk_matches_explicit_loops(shape):
  rs = np.random.RandomState(0)
  nd = no.ddpg_actor("actor", shape, True)
  P = no.init_params(nd, rs)
  for k in P:
    if k.endswith("biases"):
      P[k] = torch.tensor(rs.uniform(-0.1, 0.1, tuple(P[k].shape)))
  s = (rs.randint(0, 256, (3,) + shape).astype(np.float16) / np.float16(255))
  got = no.conv_trunk(nd, P, torch.tensor(s.astype(np.float64))).numpy()
  x = s.reshape(3, shape[0], shape[1], -1)
  want = naive_trunk(x, [P["actor/conv%d/weights" % i].numpy() for i in (1, 2, 3)],
                     [P["actor/conv%d/biases" % i].numpy() for i in (1, 2, 3)])
  assert got.shape == want.shape
  np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)


def test_naf_closed_form_a7():
  """Appendix A-7 closed forms for A=2 vs the generic L.L^T graph and autograd"""
  rs = np.random.RandomState(1)
  B = 9
  l = torch.tensor(rs.uniform(-1, 1, (B, 3)), requires_grad=True)
  mu = torch.tensor(rs.uniform(-1, 1, (B, 2)), requires_grad=True)
  V = torch.tensor(rs.uniform(-1, 1, (B, 1)), requires_grad=True)
  u = torch.tensor(rs.uniform(-1, 1, (B, 2)))
  y = torch.tensor(rs.uniform(-1, 1, (B, 1)))
  L = torch.zeros(B, 2, 2, dtype=torch.float64)
  L = torch.stack([torch.stack([torch.exp(l[:, 0]), torch.zeros(B, dtype=torch.float64)], 1),
                   torch.stack([l[:, 1], torch.exp(l[:, 2])], 1)], 1)
  d = (u - mu).unsqueeze(-1)
  A = (-0.5 * d.transpose(1, 2) @ (L @ L.transpose(1, 2)) @ d).reshape(-1, 1)
  Q = V + A
  loss = ((Q - y) ** 2).mean()
  gl, gmu, gV = torch.autograd.grad(loss, [l, mu, V])
  with torch.no_grad():
    L00, L10, L11 = torch.exp(l[:, 0]), l[:, 1], torch.exp(l[:, 2])
    d0, d1 = u[:, 0] - mu[:, 0], u[:, 1] - mu[:, 1]
    z0, z1 = L00 * d0 + L10 * d1, L11 * d1
    Acf = -0.5 * (z0 ** 2 + z1 ** 2)
    delta = 2 * (V[:, 0] + Acf - y[:, 0]) / B
    np.testing.assert_allclose(Acf.numpy(), A[:, 0].detach().numpy(), rtol=1e-12)
    np.testing.assert_allclose(gV[:, 0].numpy(), delta.numpy(), rtol=1e-12)
    np.testing.assert_allclose(gmu[:, 0].numpy(), (delta * (L00 * z0)).numpy(), rtol=1e-10)
    np.testing.assert_allclose(gmu[:, 1].numpy(), (delta * (L10 * z0 + L11 * z1)).numpy(), rtol=1e-10)
    np.testing.assert_allclose(gl[:, 0].numpy(), (-delta * d0 * z0 * L00).numpy(), rtol=1e-10)
    np.testing.assert_allclose(gl[:, 1].numpy(), (-delta * d1 * z0).numpy(), rtol=1e-10)
    np.testing.assert_allclose(gl[:, 2].numpy(), (-delta * d1 * z1 * L11).numpy(), rtol=1e-10)


def test_clip_and_optimisers():
  g = [torch.tensor([3.0, 4.0], dtype=torch.float64), torch.tensor([12.0], dtype=torch.float64)]
  gc, norm = no.clip_by_global_norm(g, 5.0)
  assert abs(float(norm) - 13.0) < 1e-12
  np.testing.assert_allclose(torch.cat(gc).numpy(), np.array([3, 4, 12]) * 5 / 13)
  gc, _ = no.clip_by_global_norm([x * 0.1 for x in g], 5.0)        # below the clip: unchanged
  np.testing.assert_allclose(torch.cat(gc).numpy(), np.array([.3, .4, 1.2]))
  # Adam, epsilon-hat placement (Appendix A-9): first step = lr*sqrt(1-b2)/(1-b1) * (1-b1) g / (sqrt((1-b2) g^2) + eps)
  p = [torch.tensor([1.0], dtype=torch.float64)]
  opt = no.Optimiser("Adam", 0.1)
  opt.apply(p, [torch.tensor([2.0], dtype=torch.float64)])
  lr_t = 0.1 * np.sqrt(1 - 0.999) / (1 - 0.9)
  want = 1.0 - lr_t * (0.1 * 2.0) / (np.sqrt(0.001 * 4.0) + 1e-8)
  assert abs(float(p[0]) - want) < 1e-12
  p = [torch.tensor([1.0], dtype=torch.float64)]
  opt = no.Optimiser("Momentum", 0.1, momentum=0.9)
  opt.apply(p, [torch.tensor([2.0], dtype=torch.float64)]); opt.apply(p, [torch.tensor([1.0], dtype=torch.float64)])
  assert abs(float(p[0]) - (1 - 0.1 * 2 - 0.1 * (0.9 * 2 + 1))) < 1e-12


def test_soft_update_association():
  t = torch.tensor([1.0, 2.0], dtype=torch.float32); s = torch.tensor([3.0, -1.0], dtype=torch.float32)
  np.testing.assert_array_equal(no.soft_update(t, s, 0.25).numpy(), (t - 0.25 * (t - s)).numpy())


def test_actor_gradient_is_batch_sum_of_dq_da():
  """Appendix A-8: the actor update direction equals -d/dtheta sum_b Q(s_b, mu_theta(s_b))"""
  rs = np.random.RandomState(3)
  shape = (2, 2, 7)
  P = {}
  a, c = no.ddpg_actor("actor", shape, False), no.ddpg_critic("critic", shape, False)
  P.update(no.init_params(a, rs)); P.update(no.init_params(c, rs))
  s1 = torch.tensor(rs.uniform(-1, 1, (6,) + shape))
  g, mu, q, dqda = no.ddpg_actor_grads(a, c, P, s1)
  leaves = no._leaf(P, no._names(a))
  obj = -no.forward(c, P, s1, no.forward(a, P, s1)).sum()
  want = torch.autograd.grad(obj, leaves)
  for x, y in zip(g, want):
    np.testing.assert_allclose(x.numpy(), y.numpy(), rtol=1e-9, atol=1e-14)


@pytest.mark.parametrize("name", ["ddpg_pixel", "ddpg_lowdim"])
def test_oracle_reproduces_committed_golden(golden_dir, name):
  g = np.load(os.path.join(golden_dir, "nets_%s.npz" % name))
  meta = json.loads(str(g["meta"]))
  P = {k[3:]: torch.tensor(g[k].astype(np.float64)) for k in g.files if k.startswith("P0/")}
  o = no.DDPGOracle(tuple(meta["state_shape"]), meta["pixels"], P)
  for step in range(2):
    batch = tuple(g["step%d/%s" % (step, f)] for f in ("s1", "a", "r", "m", "s2"))
    ra = o.actor_train(batch[0])
    np.testing.assert_allclose(torch.cat([x.reshape(-1) for x in ra["grads"]]).numpy(), g["step%d/actor_grads" % step], rtol=1e-9, atol=1e-13)
    rc = o.critic_train(batch)
    np.testing.assert_allclose(float(rc["loss"]), float(g["step%d/loss" % step]), rtol=1e-10)
    o.update_targets(0.05)
  for k, v in o.P.items():
    np.testing.assert_allclose(v.numpy(), g["Pfinal/" + k], rtol=1e-9, atol=1e-13)


def test_naf_shared_representation_sums_head_gradients():
  """--share-input-state-representation (naf_cartpole.py:151-154,176-179): V, mu and l are three `fc` heads on ONE
  representation, so d loss / d value/* carries all three heads.  Checked against central differences in fp64, and
  against the un-shared graph to make sure the flag changes the variable set the way the reference does."""
  rs = np.random.RandomState(3)
  shape, B = (3, 2, 7), 6
  value = no.naf_value("value", shape, False)
  mu, l = no.naf_shared_heads(value.fc[-2].out)
  assert [n for n, _ in mu.var_shapes()] == ["naf/output_action/fc/weights", "naf/output_action/fc/biases"]
  assert dict(l.var_shapes())["naf/l_values/fc/weights"] == (50, 3)
  P = {}
  for d in (value, mu, l):
    P.update(no.init_params(d, rs))
  P["naf/output_action/fc/weights"] = torch.tensor(rs.uniform(-0.3, 0.3, (50, 2)))
  P.update(no.retarget({k: v for k, v in P.items() if k.startswith("value/")}, "value", "target_value"))
  batch = (rs.uniform(-1, 1, (B,) + shape).astype(np.float32), rs.uniform(-1, 1, (B, 2)).astype(np.float32),
           np.ones((B, 1), np.float32), np.ones((B, 1), np.float32), rs.uniform(-1, 1, (B,) + shape).astype(np.float32))
  o = no.NAFOracle(shape, False, {k: v.clone() for k, v in P.items()}, share=True, optimiser_args={"learning_rate": 0.0})
  assert len(o.train_names) == 6 + 2 + 2
  r = o.train(batch)
  grads = dict(zip(o.train_names, r["grads"]))

  def loss_at(name, idx, eps):
    Q = {k: v.clone() for k, v in P.items()}
    Q[name].view(-1)[idx] += eps
    with torch.no_grad():
      return float(no.NAFOracle(shape, False, Q, share=True)._loss(batch)[0])
  for name in ("value/h0/weights", "value/h1/biases", "naf/l_values/fc/weights", "naf/output_action/fc/biases"):
    for idx in (0, P[name].numel() - 1):
      fd = (loss_at(name, idx, 1e-6) - loss_at(name, idx, -1e-6)) / 2e-6
      assert abs(fd - float(grads[name].reshape(-1)[idx])) <= 1e-7 * max(1.0, abs(fd)), (name, idx, fd)
  # with the heads cut off (their weights zeroed -> mu = 0, l = const) the hidden layers only see V's gradient: different
  Z = {k: v.clone() for k, v in P.items()}
  Z["naf/output_action/fc/weights"].zero_(); Z["naf/l_values/fc/weights"].zero_()
  r0 = no.NAFOracle(shape, False, Z, share=True, optimiser_args={"learning_rate": 0.0}).train(batch)
  assert not np.allclose(r0["grads"][0].numpy(), r["grads"][0].numpy())


@pytest.mark.parametrize("name", ["naf_pixel_shared", "naf_lowdim_shared"])
def test_oracle_reproduces_committed_shared_naf_golden(golden_dir, name):
  g = np.load(os.path.join(golden_dir, "nets_%s.npz" % name))
  meta = json.loads(str(g["meta"]))
  assert meta["share"] is True
  P = {k[3:]: torch.tensor(g[k].astype(np.float64)) for k in g.files if k.startswith("P0/")}
  assert not any(k.startswith("naf/output_action/h") or k.startswith("naf/l_values/conv") for k in P)
  o = no.NAFOracle(tuple(meta["state_shape"]), meta["pixels"], P, optimiser=meta["optimiser"],
                   optimiser_args=meta["optimiser_args"], share=True)
  for step in range(3):
    batch = tuple(g["step%d/%s" % (step, f)] for f in ("s1", "a", "r", "m", "s2"))
    r = o.train(batch)
    np.testing.assert_allclose(torch.cat([x.reshape(-1) for x in r["grads"]]).numpy(), g["step%d/grads" % step], rtol=1e-9, atol=1e-13)
    o.update_targets(0.05)
  for k, v in o.P.items():
    np.testing.assert_allclose(v.numpy(), g["Pfinal/" + k], rtol=1e-9, atol=1e-13)


# ------------------------------------------------------------------ --use-batch-norm (oracle side)

def _bn_net(rs, shape=(8, 8, 3, 1, 1)):
  nd = no.ddpg_actor("actor", shape, True, "8,4", batch_norm=True)
  P = no.init_params(nd, rs)
  for k in list(P):
    if k.endswith("/beta"):
      P[k] = torch.tensor(rs.uniform(-0.2, 0.2, tuple(P[k].shape)))
    if k.endswith("/moving_mean"):
      P[k] = torch.tensor(rs.uniform(-0.1, 0.1, tuple(P[k].shape)))
    if k.endswith("/moving_variance"):
      P[k] = torch.tensor(rs.uniform(0.5, 1.5, tuple(P[k].shape)))
  return nd, P


def test_batch_norm_variables_and_trainables():
  nd = no.ddpg_critic("critic", (16, 16, 3, 1, 2), True, batch_norm=True)
  names = [n for n, _ in nd.var_shapes()]
  assert names[:4] == ["critic/conv1/weights", "critic/conv1/BatchNorm/beta", "critic/conv1/BatchNorm/moving_mean",
                       "critic/conv1/BatchNorm/moving_variance"]
  assert not any(n.endswith("conv1/biases") or n.endswith("conv3/biases") for n in names)      # slim.conv2d drops the bias
  assert "critic/hidden1/biases" in names                                                      # FC layers are untouched
  tr = nd.trainable_names()
  assert "critic/conv2/BatchNorm/beta" in tr and not any("/moving_" in n for n in tr)
  assert len(names) - len(tr) == 6
  # the flag does nothing for a low-dimensional state (no conv layers)
  with_flag = [n for n, _ in no.ddpg_actor("a", (2, 2, 7), False, batch_norm=True).var_shapes()]
  assert with_flag == [n for n, _ in no.ddpg_actor("a", (2, 2, 7), False).var_shapes()]


def test_batch_norm_training_and_inference_formulas():
  rs = np.random.RandomState(5)
  x = torch.tensor(rs.randn(6, 10, 5, 4))
  beta = torch.tensor(rs.randn(10)); mm = torch.tensor(rs.randn(10) * 0.1); mv = torch.tensor(rs.uniform(0.5, 2, 10))
  xn = x.numpy()
  y = no.batch_norm(x, beta, mm, mv).numpy()                       # IS_TRAINING defaults to True, like the train ops
  mean = xn.mean(axis=(0, 2, 3)); var = xn.var(axis=(0, 2, 3))     # population variance
  want = (xn - mean[None, :, None, None]) / np.sqrt(var + 1e-3)[None, :, None, None] + beta.numpy()[None, :, None, None]
  np.testing.assert_allclose(y, want, rtol=1e-12, atol=1e-12)
  with no.is_training(False):
    y = no.batch_norm(x, beta, mm, mv).numpy()
  want = (xn - mm.numpy()[None, :, None, None]) / np.sqrt(mv.numpy() + 1e-3)[None, :, None, None] + beta.numpy()[None, :, None, None]
  np.testing.assert_allclose(y, want, rtol=1e-12, atol=1e-12)
  assert no.IS_TRAINING is True                                    # the context manager restores the flag


def test_batch_norm_gradients_flow_through_the_batch_statistics():
  """central differences through whiten -> conv -> BN(batch stats) -> ReLU -> pool x3 -> FC, for a weight, a beta and (zero
  gradient) a moving statistic"""
  rs = np.random.RandomState(6)
  nd, P = _bn_net(rs)
  s = torch.tensor(rs.uniform(0, 1, (5,) + nd.state_shape))
  co = torch.tensor(rs.randn(5, 2))

  def f(Q):
    return float((no.forward(nd, Q, s) * co).sum())
  leaves = no._leaf(P, nd.trainable_names())
  g = dict(zip(nd.trainable_names(), torch.autograd.grad((no.forward(nd, P, s) * co).sum(), leaves)))
  for name, idx in (("actor/conv1/weights", 7), ("actor/conv2/weights", 123), ("actor/conv2/BatchNorm/beta", 3), ("actor/conv3/BatchNorm/beta", 0)):
    Qp = {k: v.detach().clone() for k, v in P.items()}; Qm = {k: v.detach().clone() for k, v in P.items()}
    Qp[name].view(-1)[idx] += 1e-6; Qm[name].view(-1)[idx] -= 1e-6
    fd = (f(Qp) - f(Qm)) / 2e-6
    assert abs(fd - float(g[name].reshape(-1)[idx])) <= 2e-6 * max(1.0, abs(fd)), (name, fd, float(g[name].reshape(-1)[idx]))
  Qp = {k: v.detach().clone() for k, v in P.items()}
  Qp["actor/conv1/BatchNorm/moving_mean"] += 1.0
  assert f(Qp) == f({k: v.detach() for k, v in P.items()})        # training mode never reads the moving statistics


def test_batch_norm_train_step_semantics():
  """Appendix A-5: the moving statistics are never updated by training, are copied by the target update, and inference
  (check_loss) reads them while the train ops - target networks included - use batch statistics"""
  g = np.load(os.path.join(os.path.dirname(__file__), "golden", "nets_ddpg_pixel_bn.npz"))
  meta = json.loads(str(g["meta"]))
  assert meta["batch_norm"] is True
  P = {k[3:]: torch.tensor(g[k].astype(np.float64)) for k in g.files if k.startswith("P0/")}
  o = no.DDPGOracle(tuple(meta["state_shape"]), True, P, batch_norm=True)
  batch = tuple(g["step0/%s" % f] for f in ("s1", "a", "r", "m", "s2"))
  mm0 = o.P["critic/conv1/BatchNorm/moving_mean"].clone(); tm0 = o.P["target_critic/conv1/BatchNorm/moving_mean"].clone()
  loss_inf, _, _ = o.check_loss(batch)
  o.actor_train(batch[0]); rc = o.critic_train(batch)
  assert torch.equal(o.P["critic/conv1/BatchNorm/moving_mean"], mm0)                     # UPDATE_OPS never run
  assert abs(float(rc["loss"]) - float(loss_inf)) > 1e-6                                  # batch statistics != moving statistics
  o.update_targets(0.25)
  np.testing.assert_allclose(o.P["target_critic/conv1/BatchNorm/moving_mean"].numpy(), (tm0 - 0.25 * (tm0 - mm0)).numpy(), rtol=1e-15)
  # the TD target inside critic_train used batch statistics in the TARGET networks: recompute it by hand in training mode
  P2 = {k[3:]: torch.tensor(g[k].astype(np.float64)) for k in g.files if k.startswith("P0/")}
  o2 = no.DDPGOracle(tuple(meta["state_shape"]), True, P2, batch_norm=True)
  o2.actor_train(batch[0])
  with torch.no_grad():
    loss_train, _, _ = no.ddpg_critic_loss(o2.critic, o2.tactor, o2.tcritic, o2.P, batch, o2.discount)
  np.testing.assert_allclose(float(loss_train), float(rc["loss"]), rtol=1e-12)


def test_oracle_reproduces_committed_batch_norm_golden(golden_dir):
  g = np.load(os.path.join(golden_dir, "nets_ddpg_pixel_bn.npz"))
  meta = json.loads(str(g["meta"]))
  P = {k[3:]: torch.tensor(g[k].astype(np.float64)) for k in g.files if k.startswith("P0/")}
  o = no.DDPGOracle(tuple(meta["state_shape"]), meta["pixels"], P, batch_norm=True)
  for step in range(2):
    batch = tuple(g["step%d/%s" % (step, f)] for f in ("s1", "a", "r", "m", "s2"))
    loss0, _, _ = o.check_loss(batch)
    np.testing.assert_allclose(float(loss0), float(g["step%d/check_loss" % step]), rtol=1e-10)
    ra = o.actor_train(batch[0])
    np.testing.assert_allclose(torch.cat([x.reshape(-1) for x in ra["grads"]]).numpy(), g["step%d/actor_grads" % step], rtol=1e-9, atol=1e-13)
    rc = o.critic_train(batch)
    np.testing.assert_allclose(float(rc["loss"]), float(g["step%d/loss" % step]), rtol=1e-10)
    o.update_targets(0.05)
  for k, v in o.P.items():
    np.testing.assert_allclose(v.numpy(), g["Pfinal/" + k], rtol=1e-9, atol=1e-13)


def test_oracle_reproduces_committed_naf_batch_norm_shared_golden(golden_dir):
  g = np.load(os.path.join(golden_dir, "nets_naf_pixel_bn_shared.npz"))
  meta = json.loads(str(g["meta"]))
  assert meta["batch_norm"] is True and meta["share"] is True
  P = {k[3:]: torch.tensor(g[k].astype(np.float64)) for k in g.files if k.startswith("P0/")}
  assert "value/conv1/BatchNorm/beta" in P and "value/conv1/biases" not in P and "target_value/conv3/BatchNorm/moving_variance" in P
  o = no.NAFOracle(tuple(meta["state_shape"]), True, P, optimiser=meta["optimiser"], optimiser_args=meta["optimiser_args"],
                   share=True, batch_norm=True)
  assert not any("/moving_" in n for n in o.train_names)
  for step in range(3):
    batch = tuple(g["step%d/%s" % (step, f)] for f in ("s1", "a", "r", "m", "s2"))
    dv = o.debug_values(batch)                                      # inference mode: moving statistics
    np.testing.assert_allclose(dv[1], g["step%d/dbg_loss" % step], rtol=1e-10)
    r = o.train(batch)                                              # training mode: batch statistics, target network included
    np.testing.assert_allclose(float(r["loss"]), float(g["step%d/loss" % step]), rtol=1e-10)
    assert abs(float(r["loss"]) - float(dv[1])) > 1e-9
    np.testing.assert_allclose(torch.cat([x.reshape(-1) for x in r["grads"]]).numpy(), g["step%d/grads" % step], rtol=1e-9, atol=1e-13)
    o.update_targets(0.05)
  for k, v in o.P.items():
    np.testing.assert_allclose(v.numpy(), g["Pfinal/" + k], rtol=1e-9, atol=1e-13)


# ------------------------------------------------------------------ --use-dropout (oracle side: masks are inputs)

def test_dropout_nodes_masks_and_modes():
  rs = np.random.RandomState(11)
  shape, B = (2, 2, 7), 12
  actor = no.ddpg_actor("actor", shape, False)
  pix_critic = no.ddpg_critic("critic", (16, 16, 3, 1, 1), True)
  masks = no.draw_dropout_masks(rs, [actor, pix_critic], B)
  assert sorted(masks) == [("actor", "h0"), ("actor", "h1"), ("actor", "h2")]          # heads and the pixel critic's hidden1-3 have none
  assert masks[("actor", "h2")].shape == (B, 50) and set(np.unique(masks[("actor", "h0")].numpy())) <= {0.0, 1.0}
  P = no.init_params(actor, rs)
  P["actor/output_action/weights"] = torch.tensor(rs.uniform(-0.3, 0.3, (50, 2)))
  s = rs.uniform(-1, 1, (B,) + shape)
  plain = no.forward(actor, P, s)
  with no.dropout(masks):
    dropped = no.forward(actor, P, s)
    with no.is_training(False):
      inference = no.forward(actor, P, s)
  assert not torch.allclose(dropped, plain) and torch.equal(inference, plain)          # IS_TRAINING=False: identity
  assert no.DROPOUT_MASKS is None
  # explicit restatement of one layer chain: relu(xW+b) * mask / 0.5
  x = torch.tensor(s).reshape(B, -1)
  for i, sc in enumerate(("h0", "h1", "h2")):
    x = torch.relu(x @ P["actor/%s/weights" % sc] + P["actor/%s/biases" % sc]) * masks[("actor", sc)] * 2.0
  want = torch.tanh(x @ P["actor/output_action/weights"] + P["actor/output_action/biases"])
  np.testing.assert_allclose(dropped.numpy(), want.numpy(), rtol=1e-13, atol=1e-15)
  # all-ones masks with keep_prob 0.5 double every hidden activation; the mean over many masks is the plain activation
  h0 = torch.relu(torch.tensor(s).reshape(B, -1) @ P["actor/h0/weights"] + P["actor/h0/biases"])
  acc = torch.zeros_like(h0)
  for _ in range(400):
    acc += h0 * no.draw_dropout_masks(rs, [actor], B)[("actor", "h0")] / no.DROPOUT_KEEP_PROB
  assert float((acc / 400 - h0).abs().max()) < 0.25 * float(h0.abs().max())


def test_dropout_gradient_passes_only_through_kept_units():
  rs = np.random.RandomState(12)
  shape, B = (2, 2, 7), 5
  critic = no.ddpg_critic("critic", shape, False)
  P = no.init_params(critic, rs)
  masks = no.draw_dropout_masks(rs, [critic], B)
  s, a = rs.uniform(-1, 1, (B,) + shape), torch.tensor(rs.uniform(-1, 1, (B, 2)))
  leaves = no._leaf(P, ["critic/h2/weights", "critic/q_value/weights"])
  with no.dropout(masks):
    q = no.forward(critic, P, s, a)
    g_h2, g_q = torch.autograd.grad(q.sum(), leaves)
  # a unit of h2 dropped for EVERY sample receives no gradient; the others do
  dead = (masks[("critic", "h2")].sum(dim=0) == 0)
  if bool(dead.any()):
    assert float(g_h2[:, dead].abs().max()) == 0.0
  assert float(g_h2[:, ~dead].abs().max()) > 0.0
  # central difference on one weight of h1 under the same masks
  name, idx = "critic/h1/weights", 17
  def f(eps):
    Q = {k: v.detach().clone() for k, v in P.items()}
    Q[name].view(-1)[idx] += eps
    with no.dropout(masks):
      return float(no.forward(critic, Q, s, a).sum())
  leaves = no._leaf(P, [name])
  with no.dropout(masks):
    g, = torch.autograd.grad(no.forward(critic, P, s, a).sum(), leaves)
  fd = (f(1e-6) - f(-1e-6)) / 2e-6
  assert abs(fd - float(g.reshape(-1)[idx])) <= 1e-6 * max(1.0, abs(fd))


def test_pinned_routing_reproduces_the_graph_and_reports_disagreements():
  """oracle.gates: evaluating relu + max_pool with the routing the graph itself takes changes nothing (values, gradients);
  a routing that differs is counted and its distance from a tie reported (tests/test_gpu_step_pinned.py relies on both)"""
  import torch.nn.functional as F
  rs = np.random.RandomState(0)
  shape = (16, 16, 3, 1, 2)
  nd = no.ddpg_actor("actor", shape, True)
  P = no.init_params(nd, rs)
  s = torch.tensor(rs.rand(4, *shape))
  leaves = no._leaf(P, no._trainable(nd))
  out = no.forward(nd, P, s)
  g0 = torch.autograd.grad(out.sum(), leaves)
  routing = {}
  x = no.whiten(s.reshape(4, 16, 16, 6)).permute(0, 3, 1, 2)
  with torch.no_grad():
    for name, k in (("conv1", 5), ("conv2", 5), ("conv3", 3)):
      x = F.conv2d(x, P["actor/%s/weights" % name].permute(3, 2, 0, 1), P["actor/%s/biases" % name], padding=k // 2)
      B, Cc, H, W = x.shape
      win = x.reshape(B, Cc, H // 2, 2, W // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(B, Cc, H // 2, W // 2, 4)
      best, arg = win.max(dim=4)
      routing[("actor", name)] = torch.where(best > 0, arg, torch.full_like(arg, 4)).permute(0, 2, 3, 1).to(torch.uint8).numpy()
      x = F.max_pool2d(F.relu(x), 2)
  with no.gates(routing) as stats:
    out1 = no.forward(nd, P, s)
    g1 = torch.autograd.grad(out1.sum(), leaves)
  assert all(st["mismatched"] == 0 and st["worst_gap_rel"] == 0.0 for st in stats.values()) and len(stats) == 3
  assert float((out - out1).abs().max()) < 1e-14 and max(float((a - b).abs().max()) for a, b in zip(g0, g1)) < 1e-12
  routing[("actor", "conv1")][0, 0, 0, 0] = (routing[("actor", "conv1")][0, 0, 0, 0] + 1) % 4
  with no.gates(routing) as stats:
    no.forward(nd, P, s)
  assert stats[("actor", "conv1")]["mismatched"] == 1 and stats[("actor", "conv1")]["worst_gap_rel"] > 1e-3
  assert no.GATES is None
