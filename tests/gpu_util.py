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


FLIP = 3e-4    # allowance for max-pool argmax / ReLU gate decisions that differ between fp32 and fp64 (see below)


def assert_grads_close(got_flat, want64, want32, names, tol=TOL, slack=4.0, flip=FLIP, what="grads"):
  """Per-variable gradient parity of a whole training step at FULL size.

  Among the ~3.4 M gates (2x2 max-pool winners, ReLUs) of one c3 forward pass a handful sit within fp32 rounding of
  a tie; an fp32 run and the fp64 oracle then route one gradient term differently.  That is a discrete change, not
  a rounding error: it shows up as 1e-5..4e-4 (relative to the variable's own max) in whichever variables lie
  upstream of the flipped gate, and the reference's own fp32 CPU path shows exactly the same thing against fp64
  (SURVEY.md 7.2 'Parity definition'; e.g. actor/conv1/weights 4.4e-4 at seed 77).  The arithmetic itself is
  held to 1e-5 at full size by tests/test_gpu_kernels.py, which pins the routing, and by the small golden cases,
  which have too few gates to flip.  Here each variable must be within `tol` of fp64, or within `slack` x the fp32
  CPU path's own error, or within the flip allowance.  Returns {name: (gpu_err, cpu32_err)}."""
  got_flat = np.asarray(got_flat, dtype=np.float64).reshape(-1)
  off, rep = 0, {}
  for n, w64, w32 in zip(names, want64, want32):
    w64 = np.asarray(w64, dtype=np.float64).reshape(-1); w32 = np.asarray(w32, dtype=np.float64).reshape(-1)
    g = got_flat[off:off + w64.size]; off += w64.size
    den = max(np.abs(w64).max(), 1e-30)
    e_gpu, e_cpu = float(np.abs(g - w64).max() / den), float(np.abs(w32 - w64).max() / den)
    rep[n] = (e_gpu, e_cpu)
    if e_gpu <= max(tol, slack * e_cpu, flip):
      continue
    # a flipped conv gate with an unusually large gradient behind it: allowed for conv variables only, and only when the
    # bulk of the variable's entries still meets the arithmetic tolerance (a systematic error would touch all of them)
    frac_ok = float((np.abs(g - w64) / den <= max(tol, slack * e_cpu)).mean())
    assert "/conv" in n and frac_ok >= 0.7 and e_gpu <= 5e-2, \
        "%s %s: GPU rel err %.3e vs fp64 (fp32 CPU path: %.3e), %.0f %% of entries within tolerance" % (what, n, e_gpu, e_cpu, 100 * frac_ok)
  assert off == got_flat.size
  return rep


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


def golden_flat(g, net, prefix):
  return np.concatenate([np.asarray(g[prefix + n], dtype=np.float64).reshape(-1) for n in names_of(net)])


def set_route(tc):
  """pin conv1 to the tensor-core kernels (True) or the exact-fp32 CUDA-core kernels (False); None = default"""
  from cartpoleplusplus_b200 import _lib
  _lib.check(_lib.lib().cpp_set_option(b"conv1_tc", -1 if tc is None else int(bool(tc))))


def assert_flat_grads_close(got_flat, want_flat, cpu32_flat, nets, tc_route, what="grads"):
  """golden-vector gradient check.  CUDA-core route: the whole vector within 1e-5 (or the fp32 CPU path's own error).
  Tensor-core route: conv1 outputs carry ~1e-6 relative error instead of ~1e-7, which makes a max-pool / ReLU gate that
  sits within rounding of a tie ten times more likely to route differently from the fp64 oracle even in the small golden
  cases; there the per-variable flip-aware criterion of assert_grads_close applies."""
  got_flat = np.asarray(got_flat, dtype=np.float64).reshape(-1)
  want_flat = np.asarray(want_flat, dtype=np.float64).reshape(-1); cpu32_flat = np.asarray(cpu32_flat, dtype=np.float64).reshape(-1)
  if not tc_route:
    return assert_close(got_flat, want_flat, what=what, cpu32=cpu32_flat)
  names, w64, w32, off = [], [], [], 0
  for net in nets:
    for v in net._variables():
      n = int(np.prod(v.shape))
      names.append(v.name); w64.append(want_flat[off:off + n]); w32.append(cpu32_flat[off:off + n]); off += n
  assert off == want_flat.size, (off, want_flat.size)
  worst = 0.0
  off = 0
  for n, a64, a32 in zip(names, w64, w32):
    g = got_flat[off:off + a64.size]; off += a64.size
    den = max(np.abs(a64).max(), 1e-30)
    diff = np.abs(g - a64) / den
    e_gpu, e_cpu = float(diff.max()), float(np.abs(a32 - a64).max() / den)
    # ill-conditioned quantities (cancelling sums, e.g. the bias gradient of the 1e-3-weight tanh head) are bounded by the
    # fp32 CPU path's own error; the tensor-core route carries ~1e-6 instead of ~1.2e-7 relative forward error, hence 8x
    if e_gpu <= max(TOL, 8.0 * e_cpu):
      worst = max(worst, e_gpu)
      continue
    # Only a conv variable may exceed the arithmetic tolerance, and only in the way a flipped gate does it: the golden
    # batches are tiny (one gate is ~1/500 of a conv1 filter's gradient terms), a flipped conv-k gate touches ONE output
    # channel (<= 10 % + of the entries) of the conv layers at or below k, everything else stays within 1e-5.
    assert "/conv" in n, "%s %s: GPU rel err %.3e vs fp64 (fp32 CPU path: %.3e)" % (what, n, e_gpu, e_cpu)
    frac_ok = float((diff <= max(TOL, 8.0 * e_cpu)).mean())
    assert frac_ok >= 0.7 and e_gpu <= 5e-2, "%s %s: rel err %.3e, only %.0f %% of the entries within tolerance" % (what, n, e_gpu, 100 * frac_ok)
    worst = max(worst, e_gpu)
  return worst


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
