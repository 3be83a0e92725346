"""GPU parity of the network arithmetic (a3-a14) through the drop-in classes / C ABI against the fp64 oracle:
committed golden vectors (small), live oracle runs at BASELINE sizes, and size-independent properties.
Tolerance: 1e-5 relative (north star), measured per tensor as max|a-b|/max|b|."""
import json
import numpy as np
import pytest
import torch

from tests import gpu_util as U
from oracle import nets_oracle as no

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------ DDPG
@pytest.fixture(autouse=True)
def _default_route():
  yield
  U.set_route(None)


DDPG_PARTS = [(0, "actor"), (1, "critic")]
NAF_PARTS = [(0, "value"), (1, "naf/output_action"), (2, "naf/l_values")]


def _golden_P(g, dtype=torch.float64):
  return {k: torch.tensor(np.asarray(v), dtype=dtype) for k, v in U.golden_values(g).items()}


@pytest.mark.parametrize("tc", [False, True], ids=["cudacore", "tensorcore"])
@pytest.mark.parametrize("name", ["ddpg_pixel", "ddpg_pixel_odd", "ddpg_lowdim", "ddpg_pixel_bn"])
def test_ddpg_golden(golden_dir, name, tc):
  """committed golden vectors: forward observables against the fixture; gradients and parameters against the fp64 oracle run
  alongside from the fixture's weights and batches WITH the GPU's routing (gpu_util, 'routing-pinned gradient parity'), which
  is also held to the fixture whenever no routing decision differs"""
  U.set_route(tc)
  g, meta = U.load_golden(golden_dir, name)
  shape, pixels, B = tuple(meta["state_shape"]), meta["pixels"], meta["B"]
  bn = bool(meta.get("batch_norm", False))       # --use-batch-norm (base_network.py:74-79): exact-fp32 route only, `tc` changes nothing
  nets, eng, o = U.make_ddpg(shape, pixels, U.golden_values(g), batch_size=B, extra=["--use-batch-norm"] if bn else [])
  orc = no.DDPGOracle(shape, pixels, _golden_P(g), batch_norm=bn)
  worst, differing = {}, 0
  for step in range(2):
    batch = U.golden_batch(g, step)
    loss, td, q = nets["critic"].check_loss(batch)                        # a9: the reference's own observable
    worst["check_loss"] = U.assert_close(loss, g["step%d/check_loss" % step], what="check_loss")
    worst["check_td"] = U.assert_close(td, g["step%d/check_td" % step], what="td")
    worst["check_q"] = U.assert_close(q, g["step%d/check_q" % step], what="q")
    eng.actor_backward(batch.state_1)
    ga = eng.buffers["grads"][:eng.n_actor].cpu().numpy()
    with no.gates(U.conv_routing(eng, DDPG_PARTS, shape, B) if pixels else {}) as stats:
      ra = orc.actor_train(tuple(batch)[0])
    n_diff = U.check_gate_stats(stats)
    rep = U.per_variable_errors(U.names_of(nets["actor"]), ga, U.with_moving(nets["actor"], [x.numpy() for x in ra["grads"]]))
    worst["actor_grads"] = max(worst.get("actor_grads", 0), U.assert_all_within(rep, "step %d actor grads" % step))
    if n_diff == 0 and differing == 0:
      U.assert_close(torch.cat([x.reshape(-1) for x in ra["grads"]]).numpy(), g["step%d/actor_grads" % step], tol=1e-9, what="live oracle vs fixture")
    differing += n_diff
    eng.actor_apply()
    eng.critic_backward(batch)
    gc = eng.buffers["grads"][eng.off_critic:eng.off_critic + eng.n_critic].cpu().numpy()
    with no.gates(U.conv_routing(eng, DDPG_PARTS[1:], shape, B) if pixels else {}) as stats:
      rc = orc.critic_train(tuple(batch))
    n_diff = U.check_gate_stats(stats)
    rep = U.per_variable_errors(U.names_of(nets["critic"]), gc, U.with_moving(nets["critic"], [x.numpy() for x in rc["grads"]]))
    worst["critic_grads"] = max(worst.get("critic_grads", 0), U.assert_all_within(rep, "step %d critic grads" % step))
    if n_diff == 0 and differing == 0:
      U.assert_close(torch.cat([x.reshape(-1) for x in rc["grads"]]).numpy(), g["step%d/critic_grads" % step], tol=1e-9, what="live oracle vs fixture")
    differing += n_diff
    U.assert_close(eng.last_loss(), g["step%d/loss" % step], what="loss")
    eng.critic_apply()
    for t, s_ in (("target_actor", "actor"), ("target_critic", "critic")):
      nets[t]._run_copy_op(nets[t]._create_variables_copy_op(nets[s_], 0.05))
    orc.update_targets(0.05)
  for k, net in nets.items():
    want = np.concatenate([orc.P[n].numpy().reshape(-1) for n in U.names_of(net)])
    worst["P_" + k] = U.assert_close(U.flat_of(net), want, what="params " + k)
    if differing == 0:
      U.assert_close(want, U.golden_flat(g, net, "Pfinal/"), tol=1e-9, what="live oracle vs fixture, params " + k)
  act = nets["actor"].action_given(U.golden_batch(g, 1).state_1[0])
  assert act.shape == (1, 2)
  U.assert_close(act, orc.action_given(U.golden_batch(g, 1).state_1[0]).numpy(), what="action_given")
  worst["differing_routing_decisions"] = differing
  print(name, json.dumps(worst))


def test_ddpg_train_api_equals_backward_apply(golden_dir):
  """actor.train / critic.train (the reference methods) == backward + apply; reusing the critic trunk computed
  during actor.train gives bit-identical critic gradients"""
  g, meta = U.load_golden(golden_dir, "ddpg_pixel")
  shape = tuple(meta["state_shape"])
  batch = U.golden_batch(g, 0)
  res = []
  for mode in range(4):
    nets, eng, o = U.make_ddpg(shape, True, U.golden_values(g), batch_size=meta["B"])
    if mode == 3:
      eng.train_step(batch)               # the fused step: one backward of both networks, one apply
    elif mode == 0:
      nets["actor"].train(batch.state_1); nets["critic"].train(batch)
    elif mode == 1:
      eng.actor_backward(batch.state_1); eng.actor_apply(); eng.critic_backward(batch); eng.critic_apply()
    else:
      eng.actor_train(batch.state_1); eng.critic_train(batch, reuse_s1_trunk=True)
    res.append(torch.cat([eng.buffers["params"], eng.buffers["grads"]]).cpu().numpy())
  assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], res[2])
  n = res[0].size - 4                     # the loss/flag tail of the gradient buffer is compared separately
  U.assert_close(res[3][:n], res[0][:n], tol=1e-6, what="fused step vs actor.train; critic.train")
  U.assert_close(res[3][n], res[0][n], tol=1e-6, what="loss")


def _oracle_ddpg(shape, pixels, B, seed, dtype=torch.float64):
  from oracle.make_golden import ddpg_params, _batch
  rs = np.random.RandomState(seed)
  P = ddpg_params(rs, shape, pixels)
  batch = _batch(rs, B, shape)
  return P, batch


@pytest.mark.parametrize("shape,B", [((64, 64, 3, 1, 3), 256), ((50, 50, 3, 1, 2), 128), ((128, 128, 3, 2, 4), 12)],
                         ids=["c3", "default50", "c5shape"])
def test_ddpg_full_size_vs_live_oracle(shape, B):
  """BASELINE config 3 (64x64, R=3, C=1, batch 256), the reference's default 50x50 render and the c5 layer shapes through the
  reference's own methods (actor.train; critic.train as separate eager calls): forward observables, every gradient tensor
  and the parameters after the step within 1e-5 of the fp64 oracle evaluated with the GPU's routing
  (tests/test_gpu_step_pinned.py does the same for the fused, graph-replayed call that bench.py times)"""
  P, batch = _oracle_ddpg(shape, True, B, 77)
  values = {k: v.numpy() for k, v in P.items()}
  nets, eng, o = U.make_ddpg(shape, True, values, batch_size=B)
  orc = no.DDPGOracle(shape, True, P)
  b = U.Batch(*batch)
  l0, td0, q0 = orc.check_loss(batch)
  loss, td, q = nets["critic"].check_loss(b)
  e = dict(loss=U.assert_close(loss, l0.numpy(), what="loss"), td=U.assert_close(td, td0.numpy(), what="td"),
           q=U.assert_close(q, q0.numpy(), what="q"))
  eng.actor_backward(b.state_1)
  with no.gates(U.conv_routing(eng, DDPG_PARTS, shape, B)) as stats:
    ra = orc.actor_train(batch[0])
  U.check_gate_stats(stats, e)
  rep = U.per_variable_errors(U.names_of(nets["actor"]), eng.buffers["grads"][:eng.n_actor].cpu().numpy(), [x.numpy() for x in ra["grads"]])
  eng.actor_apply()
  eng.critic_backward(b, reuse_s1_trunk=True)
  with no.gates(U.conv_routing(eng, DDPG_PARTS[1:], shape, B)) as stats:
    rc = orc.critic_train(batch)
  U.check_gate_stats(stats, e)
  rep.update(U.per_variable_errors(U.names_of(nets["critic"]), eng.buffers["grads"][eng.off_critic:eng.off_critic + eng.n_critic].cpu().numpy(),
                                   [x.numpy() for x in rc["grads"]]))
  print("per-variable gradient errors vs fp64 (routing pinned):", json.dumps({k: "%.2e" % v for k, v in rep.items()}))
  U.assert_all_within(rep, "DDPG %s B=%d" % (shape, B))
  e["loss_step"] = U.assert_close(eng.last_loss(), float(rc["loss"]), what="loss")
  eng.critic_apply()
  for k in ("actor", "critic"):
    want = np.concatenate([orc.P[n].numpy().reshape(-1) for n in U.names_of(nets[k])])
    e["P_" + k] = U.assert_close(U.flat_of(nets[k]), want, what="params " + k)
  print("full-size errors vs fp64 oracle:", json.dumps(e))


def test_ddpg_data_parallel_linearity():
  """8e on one GPU: gradients of two half batches computed with B_global = B (and the global-batch whitening
  statistics pinned) sum to the full-batch gradients - the algebra the NCCL all-reduce relies on"""
  import ctypes as C
  from cartpoleplusplus_b200 import _lib
  shape, B = (32, 32, 3, 1, 2), 32
  P, batch = _oracle_ddpg(shape, True, B, 5)
  values = {k: v.numpy() for k, v in P.items()}
  nets, eng, o = U.make_ddpg(shape, True, values, batch_size=B)
  b = U.Batch(*[torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in batch])
  eng.actor_backward(b.state_1); eng.critic_backward(b)
  full = eng.buffers["grads"].clone()
  lib = _lib.lib()
  # global statistics of s1 / s2
  mi = []
  for s in (b.state_1, b.state_2):
    scratch = torch.zeros(int(lib.cpp_moments_scratch_doubles(6)), dtype=torch.float64, device="cuda")
    out = torch.zeros(12, dtype=torch.float32, device="cuda")
    _lib.check(lib.cpp_channel_moments(_lib.ptr(s), 1, C.c_int64(B * 32 * 32), 6, _lib.ptr(scratch), _lib.ptr(out), _lib.stream_ptr()))
    mi.append(out)
  _lib.check(lib.cpp_ddpg_set_moments(eng.handle, _lib.ptr(mi[0]), _lib.ptr(mi[1])))
  eng.world_size = 2
  acc = torch.zeros_like(full)
  for r in range(2):
    sl = slice(r * B // 2, (r + 1) * B // 2)
    hb = U.Batch(*[x[sl].contiguous() for x in b])
    eng.actor_backward(hb.state_1); eng.critic_backward(hb)
    acc += eng.buffers["grads"]
  eng.world_size = 1
  _lib.check(lib.cpp_ddpg_set_moments(eng.handle, None, None))
  U.assert_close(acc[:eng.off_loss + 1].cpu().numpy(), full[:eng.off_loss + 1].cpu().numpy(), tol=2e-6, what="sharded grads + loss")


def test_target_update_properties():
  shape = (2, 2, 7)
  nets, eng, o = U.make_ddpg(shape, False, None, batch_size=4)
  t0 = U.flat_of(nets["target_actor"]).copy(); s0 = U.flat_of(nets["actor"]).copy()
  nets["target_actor"]._run_copy_op(nets["target_actor"]._create_variables_copy_op(nets["actor"], 0.0))
  assert np.array_equal(U.flat_of(nets["target_actor"]), t0)                       # c = 0: untouched
  with pytest.raises(Exception, match="not a target network"):
    nets["actor"].update_weights()                                                  # base_network.py:47-48
  nets["target_actor"].set_as_target_network_for(nets["actor"], 0.25)
  want = t0 - np.float32(1.0) * (t0 - s0)                                            # Appendix A-11: not always == s0
  assert np.array_equal(U.flat_of(nets["target_actor"]), want)
  nets["target_actor"].update_weights()
  t1 = want
  assert np.array_equal(U.flat_of(nets["target_actor"]), t1 - np.float32(0.25) * (t1 - s0))
  assert np.array_equal(U.flat_of(nets["actor"]), s0)


# ------------------------------------------------------------------------------------------ NAF
@pytest.mark.parametrize("tc", [False, True], ids=["cudacore", "tensorcore"])
@pytest.mark.parametrize("name", ["naf_pixel", "naf_lowdim", "naf_pixel_shared", "naf_lowdim_shared", "naf_pixel_bn_shared"])
def test_naf_golden(golden_dir, name, tc):
  U.set_route(tc)
  g, meta = U.load_golden(golden_dir, name)
  shape, pixels, share, B = tuple(meta["state_shape"]), meta["pixels"], bool(meta.get("share", False)), meta["B"]
  bn = bool(meta.get("batch_norm", False))
  naf, nets, eng, o = U.make_naf(shape, pixels, U.golden_values(g), batch_size=B,
                                 optimiser=meta["optimiser"], optimiser_args=meta["optimiser_args"],
                                 extra=(["--share-input-state-representation"] if share else []) + (["--use-batch-norm"] if bn else []))
  # the fp64 oracle run alongside with the GPU's routing (the reference the gradients are held to), and the fp32 CPU path for
  # the conditioning of the Adam / Momentum parameter updates
  orc = no.NAFOracle(shape, pixels, _golden_P(g), optimiser=meta["optimiser"], optimiser_args=meta["optimiser_args"], share=share, batch_norm=bn)
  orc32 = no.NAFOracle(shape, pixels, _golden_P(g, torch.float32), optimiser=meta["optimiser"], optimiser_args=meta["optimiser_args"], share=share,
                       batch_norm=bn)
  parts = NAF_PARTS[:1] if share else NAF_PARTS
  names = U.names_of(nets["value"]) + U.names_of(nets["mu"]) + U.names_of(nets["l"])
  worst, differing = {}, 0
  for step in range(3):
    batch = U.golden_batch(g, step)
    dv = naf.debug_values(batch)
    # after an Adam / Momentum update the parameters themselves are only as close to fp64 as fp32 arithmetic allows (see the
    # final parameter check below); from step 1 on the forward values inherit that and are bounded by the fp32 CPU path's own error
    dv32 = orc32.debug_values(tuple(batch))
    for f, v, v32 in zip(("l_values", "dbg_loss", "V", "A", "V2"), dv, dv32):
      worst[f] = max(worst.get(f, 0), U.assert_close(v, g["step%d/%s" % (step, f)], what="step %d %s" % (step, f),
                                                     cpu32=v32 if step > 0 else None))
    eng.backward(batch)
    with no.gates(U.conv_routing(eng, parts, shape, B) if pixels else {}) as stats:
      r = orc.train(tuple(batch))
    n_diff = U.check_gate_stats(stats)
    r32 = orc32.train(tuple(batch))
    gr = eng.buffers["grads"].cpu().numpy()
    got = np.concatenate([gr[:eng.n_v], gr[eng.off_m:eng.off_m + eng.n_m], gr[eng.off_l:eng.off_l + eng.n_l]])
    three = [nets["value"], nets["mu"], nets["l"]]
    rep = U.per_variable_errors(names, got, U.with_moving(three, [x.numpy() for x in r["grads"]]))
    c32 = U.per_variable_errors(names, np.concatenate([np.asarray(x).reshape(-1) for x in U.with_moving(three, [x.numpy() for x in r32["grads"]])]),
                                U.with_moving(three, [x.numpy() for x in r["grads"]]))
    worst["grads"] = max(worst.get("grads", 0), U.assert_all_within(rep, "step %d grads" % step, cpu32=c32 if step > 0 else None))
    if n_diff == 0 and differing == 0:
      U.assert_close(torch.cat([x.reshape(-1) for x in r["grads"]]).numpy(), g["step%d/grads" % step], tol=1e-9, what="live oracle vs fixture")
    differing += n_diff
    loss = eng.apply(True)
    U.assert_close(loss, g["step%d/loss" % step], what="loss", cpu32=float(r32["loss"]) if step > 0 else None)
    nets["target_value"]._run_copy_op(nets["target_value"]._create_variables_copy_op(nets["value"], 0.05))
    orc.update_targets(0.05); orc32.update_targets(0.05)
  for k, net in nets.items():
    want = np.concatenate([orc.P[n].numpy().reshape(-1) for n in U.names_of(net)])
    c32 = np.concatenate([orc32.P[n].numpy().reshape(-1) for n in U.names_of(net)])
    worst["P_" + k] = U.assert_close(U.flat_of(net), want, what="params " + k, cpu32=c32)
    if differing == 0:
      U.assert_close(want, U.golden_flat(g, net, "Pfinal/"), tol=1e-9, what="live oracle vs fixture, params " + k)
  act = naf.action_given(U.golden_batch(g, 2).state_1[0], add_noise=False)
  U.assert_close(act, orc.action_given(U.golden_batch(g, 2).state_1[0]).numpy(), what="action_given",
                 cpu32=orc32.action_given(U.golden_batch(g, 2).state_1[0]).numpy())
  worst["differing_routing_decisions"] = differing
  print(name, json.dumps(worst))


@pytest.mark.parametrize("share", [False, True], ids=["three_trunks", "shared_representation"])
def test_naf_full_size_c4_shard_vs_live_oracle(share):
  """BASELINE config 4 (64x64, R=3, C=2 -> 18 channels): one 128-sample shard of the 512 batch; also with
  --share-input-state-representation (one trunk, three heads)"""
  from oracle.make_golden import _batch
  shape, B = (64, 64, 3, 2, 3), 128
  rs = np.random.RandomState(91)
  P = {}
  value = no.naf_value("value", shape, True)
  heads = no.naf_shared_heads(value.fc[-2].out) if share else (no.naf_mu(shape, True), no.naf_l(shape, True))
  for d in (value,) + tuple(heads):
    P.update(no.init_params(d, rs))
  if share:     # the reference's U(+-1e-3) action head would hide the heads' share of the trunk gradient
    P["naf/output_action/fc/weights"] = torch.tensor(rs.uniform(-0.3, 0.3, (50, 2)).astype(np.float32), dtype=torch.float64)
  P.update(no.retarget({k: v for k, v in P.items() if k.startswith("value/")}, "value", "target_value"))
  batch = _batch(rs, B, shape)
  naf, nets, eng, o = U.make_naf(shape, True, {k: v.numpy() for k, v in P.items()}, batch_size=B,
                                 optimiser="Momentum", optimiser_args={"learning_rate": 0.01, "momentum": 0.9},
                                 extra=["--share-input-state-representation"] if share else [])
  orc = no.NAFOracle(shape, True, P, optimiser="Momentum", optimiser_args={"learning_rate": 0.01, "momentum": 0.9}, share=share)
  eng.backward(U.Batch(*batch))
  e = {}
  with no.gates(U.conv_routing(eng, NAF_PARTS[:1] if share else NAF_PARTS, shape, B)) as stats:
    r = orc.train(batch)
  U.check_gate_stats(stats, e)
  gr = eng.buffers["grads"].cpu().numpy()
  got = np.concatenate([gr[:eng.n_v], gr[eng.off_m:eng.off_m + eng.n_m], gr[eng.off_l:eng.off_l + eng.n_l]])
  names = U.names_of(nets["value"]) + U.names_of(nets["mu"]) + U.names_of(nets["l"])
  rep = U.per_variable_errors(names, got, [x.numpy() for x in r["grads"]])
  print("per-variable gradient errors vs fp64 (routing pinned):", json.dumps({k: "%.2e" % v for k, v in rep.items()}))
  U.assert_all_within(rep, "NAF c4 shard share=%s" % share)
  e["loss"] = U.assert_close(eng.apply(True), float(r["loss"]), what="loss")
  for k in ("value", "mu", "l"):
    want = np.concatenate([orc.P[n].numpy().reshape(-1) for n in U.names_of(nets[k])])
    e["P_" + k] = U.assert_close(U.flat_of(nets[k]), want, what="params " + k)
  print("c4 shard errors vs fp64 oracle:", json.dumps(e))


def test_naf_check_numerics_raises():
  from cartpoleplusplus_b200 import CppError
  shape = (2, 2, 7)
  naf, nets, eng, o = U.make_naf(shape, False, None, batch_size=4)
  before = eng.buffers["params"].clone()
  rs = np.random.RandomState(0)
  s = rs.uniform(-1, 1, (4,) + shape).astype(np.float16)
  bad = U.Batch(s, np.array([[np.inf, 0]] * 4, np.float32), np.ones((4, 1), np.float32), np.ones((4, 1), np.float32), s)
  with pytest.raises(CppError) as ei:
    naf.train(bad)
  assert ei.value.status == -4                                          # CPP_ERR_NUMERICS
  assert torch.equal(eng.buffers["params"], before)                     # the step was not applied


# ------------------------------------------------------------------------------------------ LRPG
def test_lrpg_golden(golden_dir):
  from cartpoleplusplus_b200 import lrpg_cartpole
  from cartpoleplusplus_b200.synthetic_env import SyntheticCartpole
  g, meta = U.load_golden(golden_dir, "lrpg")
  o = lrpg_cartpole.set_opts(lrpg_cartpole.default_opts(["--optimiser=Adam", "--optimiser-args={\"learning_rate\": 0.01}"]))
  agent = lrpg_cartpole.LikelihoodRatioPolicyGradientAgent(SyntheticCartpole(o, discrete_actions=True))
  agent.set_variables(U.golden_values(g))
  orc32 = no.LRPGOracle((2, 2, 7), {k: torch.tensor(v, dtype=torch.float32) for k, v in U.golden_values(g).items()},
                        optimiser="Adam", optimiser_args={"learning_rate": 0.01})
  worst = {}
  for step in (0, 2):
    obs, act, adv = g["step%d/obs" % step], g["step%d/act" % step], g["step%d/adv" % step]
    worst["logits"] = U.assert_close(agent.logits_given(obs), g["step%d/logits" % step], what="logits")
    r32 = orc32.train(obs, act, adv)
    loss = agent.train(list(obs), list(act), list(adv))
    # the loss is a cancelling sum (standardised advantages have zero mean): conditioned like the fp32 CPU path
    worst["loss"] = U.assert_close(loss, g["step%d/loss" % step], what="loss", cpu32=float(r32["loss"]))
    gr = agent._engine.buffers["grads"][:agent._engine.n].cpu().numpy()
    worst["grads"] = U.assert_close(gr, g["step%d/grads" % step], what="grads", cpu32=torch.cat([x.reshape(-1) for x in r32["grads"]]).numpy())
  c32 = np.concatenate([orc32.P[n].numpy().reshape(-1) for n in U.names_of(agent)])
  worst["P"] = U.assert_close(U.flat_of(agent), U.golden_flat(g, agent, "Pfinal/"), what="params", cpu32=c32)
  print("lrpg", json.dumps(worst))


def test_empty_and_oversize_batches_are_rejected():
  from cartpoleplusplus_b200 import CppError
  nets, eng, o = U.make_ddpg((2, 2, 7), False, None, batch_size=4)
  with pytest.raises(CppError):
    eng.actor_backward(np.zeros((0, 2, 2, 7), np.float16))


def test_step_is_deterministic():
  """fixed-order reductions everywhere: two identical runs give identical bits (needed for DP replicas)"""
  P, batch = _oracle_ddpg((32, 32, 3, 2, 1), True, 16, 9)
  outs = []
  for _ in range(2):
    nets, eng, o = U.make_ddpg((32, 32, 3, 2, 1), True, {k: v.numpy() for k, v in P.items()}, batch_size=16)
    b = U.Batch(*batch)
    for _ in range(3):
      nets["actor"].train(b.state_1); nets["critic"].train(b)
    outs.append(eng.buffers["params"].cpu().numpy())
  assert np.array_equal(outs[0], outs[1])


def _set_opt(name, v):
  from cartpoleplusplus_b200 import _lib
  _lib.check(_lib.lib().cpp_set_option(name.encode(), int(v)))


def test_ddpg_fused_step_streams_and_graph_replay():
  """the fused train step on forked streams + CUDA-graph replay (eager, capture, replay, replay) gives the results of
  the single-stream eager path, and is bit-reproducible"""
  shape, B = (32, 32, 3, 1, 3), 32
  P, batch = _oracle_ddpg(shape, True, B, 21)
  values = {k: v.numpy() for k, v in P.items()}
  dev_batch = U.Batch(*[torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in batch])
  outs = []
  try:
    for streams, graphs, fused in ((0, 0, 1), (1, 1, 1), (1, 1, 1), (0, 1, 1), (1, 0, 1), (0, 0, 0), (1, 1, 0)):
      _set_opt("streams", streams); _set_opt("graphs", graphs); _set_opt("fused_mlp", fused)
      nets, eng, o = U.make_ddpg(shape, True, values, batch_size=B)
      for i in range(5):
        eng.train_step(dev_batch)
        if i == 2:
          eng.update_targets()
      torch.cuda.synchronize()
      outs.append(torch.cat([eng.buffers["params"], eng.buffers["target_params"], eng.buffers["grads"]]).cpu().numpy())
  finally:
    _set_opt("streams", -1); _set_opt("graphs", -1); _set_opt("fused_mlp", -1)
  assert np.array_equal(outs[1], outs[2])                                  # same mode twice: identical bits
  for k in (1, 3, 4, 5, 6):                                                # 5, 6: one GEMM per FC layer instead of the fused stacks
    U.assert_close(outs[k][:-4], outs[0][:-4], tol=2e-6, what="mode %d vs single-stream eager" % k)
    U.assert_close(outs[k][-4], outs[0][-4], tol=2e-6, what="loss")


def test_ddpg_schedule_switches_do_not_change_results():
  """cpp_set_option switches of the fused step's schedule (weight-prep hoisting, conv1 passes side by side or in sequence, the
  fused critic tail, SM budgets of the chains) change WHEN and WHERE work runs, not what is computed"""
  shape, B = (32, 32, 3, 1, 3), 32
  P, batch = _oracle_ddpg(shape, True, B, 22)
  values = {k: v.numpy() for k, v in P.items()}
  dev_batch = U.Batch(*[torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in batch])
  defaults = dict(prep_hoist=1, conv1_split=1, critic_tail=1, bwd_critic_sms=92, fwd_actor_sms=37)
  variants = [dict(), dict(prep_hoist=0), dict(conv1_split=0), dict(critic_tail=0), dict(bwd_critic_sms=100, fwd_actor_sms=60),
              dict(prep_hoist=0, conv1_split=0, critic_tail=0)]
  outs = []
  try:
    for var in variants:
      for k, v in dict(defaults, **var).items():
        _set_opt(k, v)
      nets, eng, o = U.make_ddpg(shape, True, values, batch_size=B)
      for i in range(4):
        eng.train_step(dev_batch)
        if i == 1:
          eng.update_targets()
      torch.cuda.synchronize()
      outs.append(torch.cat([eng.buffers["params"], eng.buffers["target_params"], eng.buffers["grads"]]).cpu().numpy())
  finally:
    for k, v in defaults.items():
      _set_opt(k, v)
  for k in range(1, len(variants)):
    U.assert_close(outs[k][:-4], outs[0][:-4], tol=2e-6, what="variant %s vs default" % variants[k])
    U.assert_close(outs[k][-4], outs[0][-4], tol=2e-6, what="loss")


# ------------------------------------------------------------------------------------------ rollout path (8f row 2)
@pytest.mark.parametrize("agent", ["ddpg", "naf", "naf_shared"])
def test_action_given_fast_path(agent):
  """ActorNetwork.action_given / NafNetwork.action_given at B = 1 (ddpg_cartpole.py:121-138, naf_cartpole.py:247-262): an fp32
  environment state made of fp16 numbers (bullet_cartpole.py:239-242) takes the tensor-core trunk through an exact fp16 copy
  and replays as one CUDA graph - same bits as feeding the fp16 array, within 1e-5 of the fp64 oracle on every replay; an fp32
  state that is NOT made of fp16 numbers is detected on the device and answered by the exact fp32 route"""
  shape = (64, 64, 3, 1, 3)
  rs = np.random.RandomState(3)
  if agent == "ddpg":
    from oracle.make_golden import ddpg_params
    P = ddpg_params(rs, shape, True)
    nets, eng, o = U.make_ddpg(shape, True, {k: v.numpy() for k, v in P.items()}, batch_size=4)
    orc = no.DDPGOracle(shape, True, P)
    act = nets["actor"].action_given
  else:
    share = agent == "naf_shared"
    value = no.naf_value("value", shape, True)
    heads = no.naf_shared_heads(value.fc[-2].out) if share else (no.naf_mu(shape, True), no.naf_l(shape, True))
    P = {}
    for d in (value,) + tuple(heads):
      P.update(no.init_params(d, rs))
    P["naf/output_action/fc/weights"] = torch.tensor(rs.uniform(-0.3, 0.3, tuple(P["naf/output_action/fc/weights"].shape)).astype(np.float32), dtype=torch.float64)
    P.update(no.retarget({k: v for k, v in P.items() if k.startswith("value/")}, "value", "target_value"))
    naf, nets, eng, o = U.make_naf(shape, True, {k: v.numpy() for k, v in P.items()}, batch_size=4,
                                   extra=["--share-input-state-representation"] if share else [])
    orc = no.NAFOracle(shape, True, P, share=share)
    act = lambda s: naf.action_given(s, add_noise=False)
  from cartpoleplusplus_b200 import _lib
  launches = []
  for trial in range(4):                                    # eager, capture, replay, replay - a new state every time
    k = rs.randint(0, 256, shape)
    s16 = (k.astype(np.float16) / np.float16(255)).astype(np.float16)
    s32 = s16.astype(np.float32)                            # what the env hands over: fp16 numbers in a float32 array
    want = orc.action_given(s16).numpy()
    l0 = _lib.lib().cpp_launch_count()
    got32 = act(s32)
    launches.append(_lib.lib().cpp_launch_count() - l0)
    assert got32.shape == (1, 2)
    U.assert_close(got32, want, what="%s action_given(fp32 env state), call %d" % (agent, trial))
    assert np.array_equal(got32, act(s16)), "fp32 env state and its fp16 copy must give the same bits"
  # an arbitrary fp32 state: flagged on the device, answered by the exact fp32 route
  s_any = rs.uniform(0, 1, shape).astype(np.float32)
  got = act(s_any)
  U.assert_close(got, orc.action_given(s_any).numpy(), what="%s action_given(arbitrary fp32 state)" % agent)
  exact = np.zeros((1, 2), np.float32)
  dev_s = torch.from_numpy(s_any[None]).cuda(); dev_o = torch.zeros(2, device="cuda")
  fn = _lib.lib().cpp_ddpg_action_given if agent == "ddpg" else _lib.lib().cpp_naf_action_given
  _lib.check(fn(eng.handle, _lib.ptr(dev_s), 0, 1, _lib.ptr(dev_o), _lib.stream_ptr()))
  assert np.array_equal(got.reshape(-1), dev_o.cpu().numpy())
  print(agent, "kernels per action_given call (eager, capture, replay, replay):", launches)
