"""--use-dropout (base_network.py:69-70: slim.dropout(keep_prob 0.5, is_training=IS_TRAINING) after every layer that
hidden_layers_starting_at creates) on the product path.  TensorFlow's random stream cannot be reproduced, so parity is checked
with INJECTED masks: the oracle takes them as inputs (oracle.nets_oracle.dropout), the library reads the same 0/1 bytes from its
mask buffers (cpp_set_option("dropout_external", 1), cpp_*_debug_view kind 3).  A second test looks at the masks the library
draws itself (Bernoulli(0.5), fresh on every call - also under CUDA-graph replay - and off outside training)."""
import ctypes as C
import json
import numpy as np
import pytest
import torch

from tests import gpu_util as U
from oracle import nets_oracle as no
from cartpoleplusplus_b200 import _lib

pytestmark = pytest.mark.gpu


def _mask_view(eng, part, layer, B):
  """torch u8 view [B][out] of the dropout mask buffer of FC layer `layer` of network `part`"""
  fn = _lib.lib().cpp_ddpg_debug_view if "actor" in eng.nets else _lib.lib().cpp_naf_debug_view
  out = (C.c_int64 * 4)()
  _lib.check(fn(eng.handle, part, 3, layer, B, out))
  off, rows, per, valid = [int(v) for v in out]
  return eng.buffers["workspace"][off:off + rows * per].view(rows, per)


def _inject(eng, part, nd, masks, B):
  for i, l in enumerate(nd.fc):
    if no.has_dropout(l.scope):
      _mask_view(eng, part, i, B).copy_(torch.from_numpy(masks[(nd.ns, l.scope)].numpy().astype(np.uint8)).cuda())


def _opt(name, v):
  _lib.check(_lib.lib().cpp_set_option(name.encode(), int(v)))


def test_ddpg_dropout_injected_masks_vs_oracle():
  """actor.train and critic.train with dropout in the actor AND the target actor (IS_TRAINING is global: ddpg_cartpole.py:237
  feeds it to the whole run); the pixel critic's hidden1-3 have none (ddpg_cartpole.py:168-171)"""
  from oracle.make_golden import ddpg_params, _batch
  shape, B = (32, 32, 3, 1, 2), 32
  rs = np.random.RandomState(5)
  P = ddpg_params(rs, shape, True)
  batch = _batch(rs, B, shape)
  nets, eng, o = U.make_ddpg(shape, True, {k: v.numpy() for k, v in P.items()}, batch_size=B, extra=["--use-dropout"])
  orc = no.DDPGOracle(shape, True, P)
  masks = no.draw_dropout_masks(rs, [orc.actor, orc.tactor], B)
  b = U.Batch(*batch)
  parts = [(0, "actor"), (1, "critic")]
  try:
    _opt("dropout_external", 1)
    eng.actor_backward(b.state_1)                      # sizes the workspace; then the masks go in and the call is repeated
    _inject(eng, 0, orc.actor, masks, B); _inject(eng, 2, orc.tactor, masks, B)
    eng.actor_backward(b.state_1)
    rep = {}
    with no.dropout(masks), no.gates(U.conv_routing(eng, parts, shape, B)) as stats:
      ra = orc.actor_train(batch[0])
    U.check_gate_stats(stats, {})
    rep.update(U.per_variable_errors(U.names_of(nets["actor"]), eng.buffers["grads"][:eng.n_actor].cpu().numpy(), [x.numpy() for x in ra["grads"]]))
    # the mask really is in the forward pass: hidden activations are zero exactly where it is
    h0 = U.debug_view(eng, 0, 2, 0, B)
    assert np.array_equal(h0 > 0, (h0 > 0) & (masks[("actor", "h0")].numpy() > 0)) and (h0 == 0).mean() > 0.45
    eng.actor_apply()
    eng.critic_backward(b)
    with no.dropout(masks), no.gates(U.conv_routing(eng, parts[1:], shape, B)) as stats:
      rc = orc.critic_train(batch)
    U.check_gate_stats(stats, {})
    rep.update(U.per_variable_errors(U.names_of(nets["critic"]), eng.buffers["grads"][eng.off_critic:eng.off_critic + eng.n_critic].cpu().numpy(),
                                     [x.numpy() for x in rc["grads"]]))
    print("dropout, injected masks: per-variable gradient errors vs fp64:", json.dumps({k: "%.2e" % v for k, v in rep.items()}))
    U.assert_all_within(rep, "DDPG --use-dropout")
    U.assert_close(eng.last_loss(), float(rc["loss"]), what="loss")
    eng.critic_apply()
    # outside training the dropout is the identity: check_loss / action_given equal the oracle without any mask
    l0, td0, q0 = orc.check_loss(batch)
    loss, td, q = nets["critic"].check_loss(b)
    U.assert_close(loss, l0.numpy(), what="check_loss"); U.assert_close(q, q0.numpy(), what="q")
    U.assert_close(nets["actor"].action_given(batch[0][0]), orc.action_given(batch[0][0]).numpy(), what="action_given")
  finally:
    _opt("dropout_external", 0)


@pytest.mark.parametrize("share", [False, True], ids=["three_trunks", "shared_representation"])
def test_naf_dropout_injected_masks_vs_oracle(share):
  """NAF: value / mu / l are all built by input_state_network(..., opts) -> dropout in every hidden layer; with
  --share-input-state-representation the two heads read the value network's dropped-out representation (ONE mask)"""
  from oracle.make_golden import _batch
  shape, B = (7, 2, 7), 32           # low-dim pose state
  rs = np.random.RandomState(6)
  P = {}
  value = no.naf_value("value", shape, False)
  heads = no.naf_shared_heads(value.fc[-2].out) if share else (no.naf_mu(shape, False), no.naf_l(shape, False))
  for d in (value,) + tuple(heads):
    P.update(no.init_params(d, rs))
  P["naf/output_action/fc/weights"] = torch.tensor(rs.uniform(-0.3, 0.3, tuple(P["naf/output_action/fc/weights"].shape)), dtype=torch.float64)
  P.update(no.retarget({k: v for k, v in P.items() if k.startswith("value/")}, "value", "target_value"))
  batch = _batch(rs, B, shape)
  naf, nets, eng, o = U.make_naf(shape, False, {k: v.numpy() for k, v in P.items()}, batch_size=B,
                                 extra=["--use-dropout"] + (["--share-input-state-representation"] if share else []))
  orc = no.NAFOracle(shape, False, P, share=share)
  nds = [orc.value, orc.tvalue] + ([] if share else [orc.mu, orc.l])
  masks = no.draw_dropout_masks(rs, nds, B)
  b = U.Batch(*batch)
  try:
    _opt("dropout_external", 1)
    eng.backward(b)
    for part, nd in [(0, orc.value), (3, orc.tvalue)] + ([] if share else [(1, orc.mu), (2, orc.l)]):
      _inject(eng, part, nd, masks, B)
    eng.backward(b)
    with no.dropout(masks):
      r = orc.train(batch)
    gr = eng.buffers["grads"].cpu().numpy()
    got = np.concatenate([gr[:eng.n_v], gr[eng.off_m:eng.off_m + eng.n_m], gr[eng.off_l:eng.off_l + eng.n_l]])
    names = U.names_of(nets["value"]) + U.names_of(nets["mu"]) + U.names_of(nets["l"])
    rep = U.per_variable_errors(names, got, [x.numpy() for x in r["grads"]])
    print("NAF dropout (share=%s), injected masks: per-variable gradient errors vs fp64:" % share, json.dumps({k: "%.2e" % v for k, v in rep.items()}))
    U.assert_all_within(rep, "NAF --use-dropout share=%s" % share)
    U.assert_close(gr[eng.off_loss], float(r["loss"]), what="loss")
    eng.apply(True)
    dv = naf.debug_values(b)                           # IS_TRAINING False: no dropout
    dvo = orc.debug_values(batch)
    for f, v, w in zip(("l_values", "loss", "V", "A", "V2"), dv, dvo):
      U.assert_close(v, np.asarray(w), what="debug_values " + f)
  finally:
    _opt("dropout_external", 0)


def test_generated_masks_are_bernoulli_half_and_fresh_under_graph_replay():
  shape, B = (32, 32, 3, 1, 2), 256
  from oracle.make_golden import ddpg_params, _batch
  rs = np.random.RandomState(7)
  P = ddpg_params(rs, shape, True)
  nets, eng, o = U.make_ddpg(shape, True, {k: v.numpy() for k, v in P.items()}, batch_size=B, extra=["--use-dropout"])
  db = U.Batch(*[torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in _batch(rs, B, shape)])
  _opt("dropout_seed", 1234)
  seen = []
  for i in range(5):                                   # eager, capture, replays of the fused step
    eng.train_step(db)
    torch.cuda.synchronize()
    seen.append([_mask_view(eng, part, layer, B).cpu().numpy().copy() for part in (0, 2) for layer in (0, 1, 2)])
  for ms in seen:
    for m in ms:
      assert set(np.unique(m)) <= {0, 1} and abs(m.mean() - 0.5) < 0.02, m.mean()
  for i in range(1, 5):                                # a new draw on every call, graph replays included
    assert all((a != b_).mean() > 0.4 for a, b_ in zip(seen[i - 1], seen[i]))
  assert (seen[0][0] != seen[0][3]).mean() > 0.4       # actor and target actor draw different masks
