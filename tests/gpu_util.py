"""Helpers shared by the GPU parity tests: build the drop-in agents with injected weights."""
import json
import os
import numpy as np
import torch

from cartpoleplusplus_b200 import base_network, ddpg_cartpole, naf_cartpole, lrpg_cartpole
from cartpoleplusplus_b200.replay_memory import Batch

TOL = 1e-5      # north star: outputs within 1e-5 rel of the CPU reference


def rel_err(got, want):
  """per-tensor max|a-b| / max|b| (SURVEY.md 7.2 'Parity definition')"""
  got = np.asarray(got, dtype=np.float64).reshape(-1)
  want = np.asarray(want, dtype=np.float64).reshape(-1)
  assert got.shape == want.shape, (got.shape, want.shape)
  den = max(np.abs(want).max(), 1e-30)
  return float(np.abs(got - want).max() / den)


def assert_close(got, want, tol=TOL, what="", cpu32=None, slack=4.0):
  """|got - want|_max / |want|_max <= tol.  cpu32: the same quantity from the fp32 CPU oracle path; when given, the
  bound is max(tol, slack * its own error against the fp64 oracle) - used only for ill-conditioned quantities
  (Adam-normalised parameter updates, a loss that is a cancelling sum) where fp32 itself cannot hold 1e-5."""
  e = rel_err(got, want)
  bound = tol if cpu32 is None else max(tol, slack * rel_err(cpu32, want))
  assert e <= bound, "%s: rel err %.3e > %.1e%s" % (what, e, bound, "" if cpu32 is None else " (fp32 CPU path: %.3e)" % rel_err(cpu32, want))
  return e


# ---- routing-pinned gradient parity ---------------------------------------------------------------------------------
# relu + 2x2 max-pool is piecewise linear; which piece applies is a discrete decision per window (winner, ReLU open or
# closed).  A decision within rounding of a tie may be taken differently by the GPU forward pass and by the fp64 oracle;
# the gradient behind it then differs by a whole term.  Instead of allowing for that, the tests read the GPU's decisions
# back (cpp_*_debug_view) and evaluate the fp64 oracle with THOSE (oracle.nets_oracle.gates): both sides then compute
# the same smooth function and every gradient tensor must meet the plain tolerance.  The decisions that differ from the
# oracle's own are counted and each must sit within rounding of a tie.
TIE = 1e-5             # a differing decision must be this close to a tie (relative to max|pre-activation| of the layer)
MAX_MISMATCH = 2e-4    # ... and differing decisions must be rare (fraction of a layer's decisions), at least allow one


def debug_view(eng, part, kind, index, B):
  """intermediate `kind` (0 routing bytes, 1 pooled conv output, 2 FC output) of layer `index` of network `part` after the
  last step, read from the engine's workspace through cpp_{ddpg,naf}_debug_view"""
  import ctypes as C
  from cartpoleplusplus_b200 import _lib
  fn = _lib.lib().cpp_ddpg_debug_view if "actor" in eng.nets else _lib.lib().cpp_naf_debug_view
  out = (C.c_int64 * 4)()
  _lib.check(fn(eng.handle, part, kind, index, B, out))
  off, rows, per, valid = [int(v) for v in out]
  ws = eng.buffers["workspace"]
  if kind == 0:
    return ws[off:off + rows * per].view(rows, per).cpu().numpy()
  return ws[off:off + rows * per * 4].view(torch.float32).view(rows, per)[:, :valid].cpu().numpy()


def conv_routing(eng, parts, shape, B):
  """{(namespace, 'conv1'|'conv2'|'conv3'): u8 [B][PH][PW][10]} for parts = [(part index, namespace), ...]"""
  out = {}
  for part, ns in parts:
    H, W = shape[0], shape[1]
    for i, name in enumerate(("conv1", "conv2", "conv3")):
      H, W = H // 2, W // 2
      out[(ns, name)] = debug_view(eng, part, 0, i, B).reshape(B, H, W, 10)
  return out


def check_gate_stats(stats, report=None):
  """every routing decision that differs from the fp64 graph's own is a tie within rounding, and there are few; -> #differing"""
  total = 0
  for key, st in sorted(stats.items()):
    total += st["mismatched"]
    if report is not None:
      report["gates %s/%s" % key] = "%d of %d differ, worst gap %.1e" % (st["mismatched"], st["n"], st["worst_gap_rel"])
    assert st["mismatched"] <= max(1, MAX_MISMATCH * st["n"]), "%s: %d of %d routing decisions differ from the fp64 graph" % (key, st["mismatched"], st["n"])
    assert st["worst_gap_rel"] <= TIE, "%s: a differing decision is %.2e away from a tie - not a rounding tie" % (key, st["worst_gap_rel"])
  return total


def per_variable_errors(names, got_flat, want_list):
  """{variable: max|got - want| / max|want|} over the flat gradient (or parameter) vector of one network"""
  got_flat = np.asarray(got_flat, dtype=np.float64).reshape(-1)
  off, rep = 0, {}
  for n, w in zip(names, want_list):
    w = np.asarray(w, dtype=np.float64).reshape(-1)
    rep[n] = rel_err(got_flat[off:off + w.size], w)
    off += w.size
  assert off == got_flat.size, (off, got_flat.size)
  return rep


def assert_all_within(rep, what, tol=TOL, cpu32=None, slack=4.0):
  """every tensor of `rep` within tol (or, for ill-conditioned quantities, slack x the fp32 CPU path's own error cpu32[name])"""
  bad = {k: "%.2e" % v for k, v in rep.items() if not v <= (tol if cpu32 is None else max(tol, slack * cpu32.get(k, 0.0)))}
  assert not bad, "%s: tensors outside %.0e of the fp64 oracle (routing pinned): %s" % (what, tol, json.dumps(bad))
  return max(rep.values()) if rep else 0.0


def load_golden(golden_dir, name):
  g = np.load(os.path.join(golden_dir, "nets_%s.npz" % name))
  return g, json.loads(str(g["meta"]))


def pixel_flags(shape):
  H, W, _, C, R = shape
  return ["--use-raw-pixels", "--render-height=%d" % H, "--render-width=%d" % W, "--num-cameras=%d" % C,
          "--action-repeats=%d" % R]


def make_ddpg(shape, pixels, values=None, batch_size=8, extra=()):
  argv = ["--batch-size=%d" % batch_size] + (pixel_flags(shape) if pixels else ["--action-repeats=%d" % shape[0]]) + list(extra)
  o = ddpg_cartpole.set_opts(ddpg_cartpole.default_opts(argv))
  s1, s2 = base_network.Placeholder(shape, "s1"), base_network.Placeholder(shape, "s2")
  actor = ddpg_cartpole.ActorNetwork("actor", s1, 2)
  critic = ddpg_cartpole.CriticNetwork("critic", actor)
  tactor = ddpg_cartpole.ActorNetwork("target_actor", s2, 2)
  tcritic = ddpg_cartpole.CriticNetwork("target_critic", tactor)
  actor.init_ops_for_training(critic)
  critic.init_ops_for_training(tcritic)
  nets = dict(actor=actor, critic=critic, target_actor=tactor, target_critic=tcritic)
  if values is not None:
    for n in nets.values():
      n.set_variables(values)
  return nets, actor._engine, o


def make_naf(shape, pixels, values=None, batch_size=8, optimiser="GradientDescent", optimiser_args=None, extra=()):
  argv = ["--batch-size=%d" % batch_size, "--optimiser=%s" % optimiser,
          "--optimiser-args=%s" % json.dumps(optimiser_args or {"learning_rate": 0.001})]
  argv += (pixel_flags(shape) if pixels else ["--action-repeats=%d" % shape[0]]) + list(extra)
  o = naf_cartpole.set_opts(naf_cartpole.default_opts(argv))
  s1, s2 = base_network.Placeholder(shape, "s1"), base_network.Placeholder(shape, "s2")
  value = naf_cartpole.ValueNetwork("value", s1, o.hidden_layers)
  tvalue = naf_cartpole.ValueNetwork("target_value", s2, o.hidden_layers)
  naf = naf_cartpole.NafNetwork("naf", s1, s2, value, tvalue, 2)
  nets = dict(value=value, target_value=tvalue, mu=naf.mu_net, l=naf.l_net)
  if values is not None:
    for n in nets.values():
      n.set_variables(values)
  return naf, nets, naf._engine, o


def golden_values(g, prefix="P0/"):
  return {k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)}


def golden_batch(g, step):
  return Batch(g["step%d/s1" % step], g["step%d/a" % step], g["step%d/r" % step], g["step%d/m" % step], g["step%d/s2" % step])


def flat_of(net):
  return net.flat_params().detach().cpu().numpy()


def names_of(net):
  return [v.name for v in net._variables()]


def with_moving(net, trainable_list):
  """the oracle's gradient list covers tf.trainable_variables; the library's flat gradient buffer has a (zero) slot for every
  variable, the batch-norm moving statistics included: pad the list with zeros at their places"""
  out, it = [], iter(trainable_list)
  for v in [v for n in (net if isinstance(net, (list, tuple)) else [net]) for v in n._variables()]:
    out.append(np.zeros(v.shape) if "/moving_" in v.name else next(it))
  assert next(it, None) is None
  return out


def golden_flat(g, net, prefix):
  return np.concatenate([np.asarray(g[prefix + n], dtype=np.float64).reshape(-1) for n in names_of(net)])


def set_route(tc):
  """pin conv1 to the tensor-core kernels (True) or the exact-fp32 CUDA-core kernels (False); None = default"""
  from cartpoleplusplus_b200 import _lib
  _lib.check(_lib.lib().cpp_set_option(b"conv1_tc", -1 if tc is None else int(bool(tc))))


def to_c24(x32):
  """fp32 activation (..., 10) -> the 24-channel fp16 piece layout of csrc/conv_tc.cuh:
  [hi0..7 | lo0..7 | hi8 hi9 lo8 lo9 1.0 0 0 0]; returns (pieces fp16 (..., 24), exact fp64 value hi + lo)"""
  x32 = np.asarray(x32, dtype=np.float32)
  assert x32.shape[-1] == 10
  hi = x32.astype(np.float16); lo = (x32 - hi.astype(np.float32)).astype(np.float16)
  out = np.zeros(x32.shape[:-1] + (24,), dtype=np.float16)
  out[..., 0:8] = hi[..., 0:8]; out[..., 8:16] = lo[..., 0:8]
  out[..., 16:18] = hi[..., 8:10]; out[..., 18:20] = lo[..., 8:10]
  out[..., 20] = 1.0
  return out, hi.astype(np.float64) + lo.astype(np.float64)


def from_c24(p):
  """(hi, lo) fp64 (..., 10) of a 24-channel piece tensor, plus the constant channel and the padding"""
  p = np.asarray(p).astype(np.float64)
  hi = np.concatenate([p[..., 0:8], p[..., 16:18]], axis=-1); lo = np.concatenate([p[..., 8:16], p[..., 18:20]], axis=-1)
  return hi, lo, p[..., 20], p[..., 21:24]
