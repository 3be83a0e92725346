"""Whole-step parity on the call bench.py times, with the routing pinned.

`eng.train_step(batch, moments=...)` (DDPG: forked streams + CUDA-graph replay, tensor-core route) and `eng.backward`
(NAF) at the BASELINE sizes are compared with the fp64 oracle evaluated with the GPU's OWN routing decisions (2x2 max-pool
winners and ReLU gates of every conv layer, read back through cpp_*_debug_view).  relu/max-pool is piecewise linear:
once the pieces are pinned, GPU and oracle compute the same smooth function and EVERY gradient tensor has to meet the
1e-5 budget - no allowance for "flipped gates".  Separately the test counts the decisions on which the GPU and the
unpinned fp64 graph disagree and checks that each of them sits within rounding of a tie (the value passed on differs by
less than 1e-5 of the layer's scale), i.e. that a disagreement is never an arithmetic error in disguise.

Reference: ddpg_cartpole.py:331-337 (actor.train; critic.train), naf_cartpole.py:367-373, base_network.py:103-123."""
import ctypes as C
import json
import numpy as np
import pytest
import torch

from tests import gpu_util as U
from oracle import nets_oracle as no
from cartpoleplusplus_b200 import _lib

pytestmark = pytest.mark.gpu

DDPG_PARTS = [(0, "actor"), (1, "critic")]
NAF_PARTS = [(0, "value"), (1, "naf/output_action"), (2, "naf/l_values")]


def _moments(lib, s, n_pix, Cin):
  scratch = torch.zeros(int(lib.cpp_moments_scratch_doubles(Cin)), dtype=torch.float64, device="cuda")
  out = torch.zeros(2 * Cin, dtype=torch.float32, device="cuda")
  _lib.check(lib.cpp_channel_moments(_lib.ptr(s), 1, C.c_int64(n_pix), Cin, _lib.ptr(scratch), _lib.ptr(out), _lib.stream_ptr()))
  return out


@pytest.mark.parametrize("shape,B", [((64, 64, 3, 1, 3), 256), ((128, 128, 3, 2, 4), 16), ((50, 50, 3, 1, 2), 64)],
                         ids=["c3", "c5shape", "default50"])
def test_ddpg_train_step_graph_replay_vs_oracle_with_pinned_routing(shape, B):
  report, rep = run_ddpg_pinned(shape, B)
  print("pinned-routing whole step %s B=%d:" % (shape, B), json.dumps(report))
  print("per-variable gradient errors vs fp64 (pinned routing):", json.dumps({k: "%.2e" % v for k, v in rep.items()}))
  U.assert_all_within(rep, "DDPG %s B=%d" % (shape, B))


def run_ddpg_pinned(shape, B, seed=77, bn=False):
  from oracle.make_golden import ddpg_params, _batch
  lib = _lib.lib()
  Cin = int(np.prod(shape[2:]))
  rs = np.random.RandomState(seed)
  P = ddpg_params(rs, shape, True, batch_norm=bn) if bn else ddpg_params(rs, shape, True)
  batch = _batch(rs, B, shape)
  nets, eng, o = U.make_ddpg(shape, True, {k: v.numpy() for k, v in P.items()}, batch_size=B, extra=["--use-batch-norm"] if bn else [])
  db = U.Batch(*[torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in batch])
  moms = (_moments(lib, db.state_1, B * shape[0] * shape[1], Cin), _moments(lib, db.state_2, B * shape[0] * shape[1], Cin))
  p0, t0 = eng.buffers["params"].clone(), eng.buffers["target_params"].clone()
  # eager run, graph capture, first replay, second replay - every time from the same parameters; the last one is checked
  results = []
  for i in range(4):
    eng.buffers["params"].copy_(p0); eng.buffers["target_params"].copy_(t0)
    eng.train_step(db, moments=moms)
    torch.cuda.synchronize()
    results.append(torch.cat([eng.buffers["params"], eng.buffers["grads"]]).cpu().numpy())
  assert np.array_equal(results[2], results[3]), "two graph replays of the same step differ"
  U.assert_close(results[3][:-4], results[0][:-4], tol=2e-6, what="graph replay vs eager run")
  routing = U.conv_routing(eng, DDPG_PARTS, shape, B)
  grads = eng.buffers["grads"].cpu().numpy()
  report = {}
  orc = no.DDPGOracle(shape, True, P, batch_norm=bn)
  with no.gates(routing) as stats:
    ra = orc.actor_train(batch[0])                       # updates orc.P[actor/*]; the critic step below does not read them
    rc = orc.critic_train(batch)
  U.check_gate_stats(stats, report)
  rep = U.per_variable_errors(U.names_of(nets["actor"]), grads[:eng.n_actor], U.with_moving(nets["actor"], [x.numpy() for x in ra["grads"]]))
  rep.update(U.per_variable_errors(U.names_of(nets["critic"]), grads[eng.off_critic:eng.off_critic + eng.n_critic],
                                   U.with_moving(nets["critic"], [x.numpy() for x in rc["grads"]])))
  report["loss"] = U.assert_close(grads[eng.off_loss], float(rc["loss"]), what="loss")
  for k in ("actor", "critic"):
    want = np.concatenate([orc.P[n].numpy().reshape(-1) for n in U.names_of(nets[k])])
    report["P_" + k] = U.assert_close(U.flat_of(nets[k]), want, what="params after the step, " + k)
  report["worst grad"] = max(rep.values())
  return report, rep


@pytest.mark.parametrize("share", [False, True], ids=["three_trunks", "shared_representation"])
def test_naf_backward_graph_replay_vs_oracle_with_pinned_routing(share):
  """BASELINE config 4 (64x64, R=3, C=2 -> 18 channels): one 128-sample shard of the 512 batch"""
  from oracle.make_golden import _batch
  lib = _lib.lib()
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
                                 extra=["--share-input-state-representation"] if share else [])
  db = U.Batch(*[torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in batch])
  results = []
  for i in range(4):                                     # eager, capture, replay, replay (no apply in between: same parameters)
    eng.backward(db)
    torch.cuda.synchronize()
    results.append(eng.buffers["grads"].cpu().numpy())
  assert np.array_equal(results[2], results[3]), "two graph replays of the same step differ"
  U.assert_close(results[3][:-4], results[0][:-4], tol=2e-6, what="graph replay vs eager run")
  routing = U.conv_routing(eng, NAF_PARTS[:1] if share else NAF_PARTS, shape, B)
  gr = results[3]
  got = np.concatenate([gr[:eng.n_v], gr[eng.off_m:eng.off_m + eng.n_m], gr[eng.off_l:eng.off_l + eng.n_l]])
  orc = no.NAFOracle(shape, True, P, share=share)
  report = {}
  with no.gates(routing) as stats:
    r = orc.train(batch)
  U.check_gate_stats(stats, report)
  names = U.names_of(nets["value"]) + U.names_of(nets["mu"]) + U.names_of(nets["l"])
  rep = U.per_variable_errors(names, got, [x.numpy() for x in r["grads"]])
  report["loss"] = U.assert_close(gr[eng.off_loss], float(r["loss"]), what="loss")
  report["worst grad"] = max(rep.values())
  print("pinned-routing NAF c4 shard (share=%s):" % share, json.dumps(report))
  print("per-variable gradient errors vs fp64 (pinned routing):", json.dumps({k: "%.2e" % v for k, v in rep.items()}))
  U.assert_all_within(rep, "NAF c4 shard share=%s" % share)


@pytest.mark.parametrize("conv_row,wgrad_tc", [(0, 1), (5, 5), (1, 3), (3, 0)],
                         ids=["round4_kernels", "separate_unpool", "piece_mode_wgrad", "cp_async_strips_mma_sync_wgrad"])
def test_ddpg_whole_step_on_the_alternative_conv_routes(conv_row, wgrad_tc):
  """The routes the row-sweep kernels replaced stay selectable (`conv_row`, `wgrad_tc`: include/cartpolepp.h) and take the shapes
  those kernels do not cover; inside the fused, graph-replayed c3 step each combination has to hold the same 1e-5 bound against
  the fp64 oracle with the routing pinned."""
  lib = _lib.lib()
  try:
    _lib.check(lib.cpp_set_option(b"conv_row", conv_row))
    _lib.check(lib.cpp_set_option(b"wgrad_tc", wgrad_tc))
    report, rep = run_ddpg_pinned((64, 64, 3, 1, 3), 64, seed=85)
    print("conv_row=%d wgrad_tc=%d whole step:" % (conv_row, wgrad_tc), json.dumps(report))
    U.assert_all_within(rep, "DDPG c3 shape, conv_row=%d wgrad_tc=%d" % (conv_row, wgrad_tc))
  finally:
    _lib.check(lib.cpp_set_option(b"conv_row", 1))
    _lib.check(lib.cpp_set_option(b"wgrad_tc", 5))


@pytest.mark.parametrize("fused_mlp", [0, 1], ids=["per_layer_forward", "fused_forward"])
def test_fc_on_tensor_cores_whole_step(fused_mlp):
  """cpp_set_option("fc_tc", 15): every fully connected GEMM with M >= 64 rows - forward (when the stacks run per layer),
  input gradients, weight gradients - goes through the tcgen05 kernel of fc_tc.cu (three bf16 pieces per fp32 operand).  The
  whole c3 step, graph-replayed, has to hold the same 1e-5 bound against the fp64 oracle with the routing pinned."""
  lib = _lib.lib()
  try:
    _lib.check(lib.cpp_set_option(b"fc_tc", 15))
    _lib.check(lib.cpp_set_option(b"fused_mlp", fused_mlp))
    n0 = int(lib.cpp_launch_count())
    report, rep = run_ddpg_pinned((64, 64, 3, 1, 3), 256, seed=81)
    print("fc_tc=15 fused_mlp=%d whole step:" % fused_mlp, json.dumps(report), "launches", int(lib.cpp_launch_count()) - n0)
    print("per-variable gradient errors vs fp64 (pinned routing):", json.dumps({k: "%.2e" % v for k, v in rep.items()}))
    U.assert_all_within(rep, "DDPG c3 with the FC layers on tcgen05")
  finally:
    _lib.check(lib.cpp_set_option(b"fc_tc", 0))
    _lib.check(lib.cpp_set_option(b"fused_mlp", -1))


@pytest.mark.parametrize("shape,B", [((64, 64, 3, 1, 3), 64), ((50, 50, 3, 1, 2), 32)], ids=["c3shape", "default50"])
def test_ddpg_batch_norm_train_step_vs_oracle_with_pinned_routing(shape, B):
  """--use-batch-norm (base_network.py:74-79; nearly every pixel experiment of the reference, exps/run_8*.sh): the same fused,
  graph-replayed call with slim.batch_norm in every conv layer of all four networks (batch statistics also in the targets,
  SURVEY.md Appendix A-5): every gradient tensor - BatchNorm/beta included, moving statistics exactly zero - within 1e-5 of
  the fp64 oracle with the routing pinned"""
  import time
  report, rep = run_ddpg_pinned(shape, B, seed=83, bn=True)
  print("pinned-routing whole step with batch norm %s B=%d:" % (shape, B), json.dumps(report))
  print("per-variable gradient errors vs fp64 (pinned routing):", json.dumps({k: "%.2e" % v for k, v in rep.items()}))
  U.assert_all_within(rep, "DDPG + batch norm %s B=%d" % (shape, B))
  assert all(v == 0.0 for k, v in rep.items() if "/moving_" in k), "moving statistics must not receive a gradient"
