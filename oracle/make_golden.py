#!/usr/bin/env python
"""Generates tests/golden/* .  Run in the BUILD container only (needs /root/reference):

    python -m oracle.make_golden

1. replay_*.npz   - outputs of the REAL reference /root/reference/replay_memory.py.  The file is
                    Python 2; it is read as text, its print statements are rewritten in memory
                    (nothing is copied into this repo) and it is exec'd with empty stubs for the
                    modules it imports but does not use on this path (tensorflow, event_log, util).
2. mt19937_kat.json - numpy legacy RandomState.randint outputs (the dependency the reference calls).
3. nets_*.npz     - fp64 oracle outputs (oracle/nets_oracle.py) on seeded weights/batches.  These are
                    NOT reference outputs (TensorFlow is unavailable): "parity unpinned", they freeze
                    the oracle so GPU parity tests have committed vectors.
"""
import json
import os
import re
import sys
import types
import numpy as np
import torch

from oracle import nets_oracle as no

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF = "/root/reference/replay_memory.py"


def load_reference_replay():
  src = open(REF).read()
  src = src.split('if __name__ == "__main__":')[0]
  out = []
  for line in src.split("\n"):
    m = re.match(r"^(\s*)print\s+>>sys\.stderr,\s*(.*)$", line)
    if m:
      line = "%sprint(%s, file=sys.stderr)" % (m.group(1), m.group(2)); out.append(line); continue
    m = re.match(r"^(\s*)print\s+(.*?)(,?)\s*$", line)
    if m and not line.strip().startswith("#"):
      line = "%sprint(%s)" % (m.group(1), m.group(2))
    out.append(line)
  for name in ("tensorflow", "event_log", "util"):
    sys.modules.setdefault(name, types.ModuleType(name))
  mod = types.ModuleType("reference_replay_memory")
  exec(compile("\n".join(out), REF, "exec"), mod.__dict__)
  return mod


def episode_stream(rs, state_shape, n_episodes, min_len, max_len):
  """deterministic synthetic episodes: states are fp16-exact k/255 values"""
  eps = []
  for _ in range(n_episodes):
    L = int(rs.randint(min_len, max_len + 1))
    def st():
      return (rs.randint(0, 256, state_shape).astype(np.float16) / np.float16(255))
    init = st()
    seq = [(rs.uniform(-1, 1, (1, 2)).astype(np.float32), float(rs.randint(0, 5)), st()) for _ in range(L)]
    eps.append((init, seq))
  return eps


def golden_replay(ref):
  cases = {
      "small":  dict(buffer_size=7, state_shape=(2, 3), load_factor=2.0, n_episodes=9, min_len=2, max_len=5, seed=11, batch=5),
      "ragged": dict(buffer_size=43, state_shape=(2, 2, 7), load_factor=1.5, n_episodes=40, min_len=3, max_len=9, seed=12, batch=13),
      "pixels": dict(buffer_size=64, state_shape=(8, 8, 3, 1, 2), load_factor=2.0, n_episodes=30, min_len=3, max_len=8, seed=13, batch=16),
  }
  for name, c in cases.items():
    rs = np.random.RandomState(c["seed"])
    eps = episode_stream(rs, c["state_shape"], c["n_episodes"], c["min_len"], c["max_len"])
    rm = ref.ReplayMemory(c["buffer_size"], c["state_shape"], 2, c["load_factor"])
    out = {"meta": json.dumps(c)}
    np.random.seed(1000 + c["seed"])
    for e, (init, seq) in enumerate(eps):
      rm.add_episode(init, seq)
      # snapshot after every episode: tables + a sampled batch from the global np.random stream
      b = rm.batch(c["batch"])
      out["e%d_insert_full_size" % e] = np.array([rm.insert, int(rm.full), rm.size(), len(rm.state_free_slots)])
      n = rm.size()
      out["e%d_s1idx" % e] = rm.state_1_idx[:n].copy()
      out["e%d_s2idx" % e] = rm.state_2_idx[:n].copy()
      out["e%d_free" % e] = np.array(rm.state_free_slots, dtype=np.int32)
      for f, v in zip(b._fields, b):
        out["e%d_batch_%s" % (e, f)] = v
    out["final_action"] = rm.action[:rm.size()].copy()
    out["final_reward"] = rm.reward[:rm.size()].copy()
    out["final_mask"] = rm.terminal_mask[:rm.size()].copy()
    # the episodes themselves, so the test can replay them without the generator
    for e, (init, seq) in enumerate(eps):
      out["ep%d_init" % e] = init
      out["ep%d_actions" % e] = np.stack([a for a, _, _ in seq])
      out["ep%d_rewards" % e] = np.array([r for _, r, _ in seq], dtype=np.float32)
      out["ep%d_states" % e] = np.stack([s for _, _, s in seq])
    np.savez_compressed(os.path.join(GOLD, "replay_%s.npz" % name), **out)
    print("replay_%s: %d episodes" % (name, len(eps)))

  # the reference's own unit-test expectations (replay_memory_test.py:19-86), re-verified on the
  # current numpy ReplayMemory
  rm = ref.ReplayMemory(buffer_size=3, state_shape=(2, 3), action_dim=2, load_factor=2)
  assert rm.size() == 0 and rm.random_indexes() == [] and all(len(x) == 0 for x in rm.batch(4))
  def s_for(i):
    return (np.array(range(1, 7)) + (10 * i)).reshape(2, 3)
  rm.add_episode(s_for(0), [((i * 10) + 7, (i * 10) + 8, s_for(i)) for i in range(1, 5)])
  rm.add_episode(s_for(5), [((i * 10) + 7, (i * 10) + 8, s_for(i)) for i in range(6, 9)])
  assert rm.size() == 3
  assert (rm.reward == [[88], [68], [78]]).all() and (rm.terminal_mask == [[0], [1], [1]]).all()
  print("reference unit-test expectations hold on the current reference code")


def golden_mt():
  kat = []
  for seed, high, n, calls in [(0, 22000, 256, 2), (42, 22000, 256, 1), (123, 200000, 1024, 1), (7, 3, 100, 3),
                               (1, 1000, 512, 1), (5, 12000, 512, 2), (9, 1, 5, 2), (3, 2, 64, 1),
                               (2 ** 31 + 5, 4096, 300, 1), (17, 65536, 700, 1), (18, 65537, 700, 1)]:
    rs = np.random.RandomState(seed % (2 ** 32))
    outs = [rs.randint(0, high, n).tolist() for _ in range(calls)]
    kat.append(dict(seed=seed % (2 ** 32), high=high, n=n, outs=outs))
  # stream interleaved with other legacy draws, as the reference's rollout loop does
  rs = np.random.RandomState(77)
  seq = []
  for i in range(4):
    seq.append(("randn", rs.randn(2).tolist()))
    seq.append(("randint", rs.randint(0, 22000, 128).tolist()))
    seq.append(("random", float(rs.random_sample())))
  json.dump(dict(kat=kat, interleaved=dict(seed=77, seq=seq)), open(os.path.join(GOLD, "mt19937_kat.json"), "w"))
  print("mt19937_kat.json: %d cases" % len(kat))


def _batch(rs, B, state_shape, sparse=False):
  def states():
    if sparse:   # constant background with a few coloured pixels: stresses whitening
      k = np.full((B,) + state_shape, 200, dtype=np.int64)
      m = rs.rand(*k.shape) < 0.03
      k[m] = rs.randint(0, 256, int(m.sum()))
    else:
      k = rs.randint(0, 256, (B,) + state_shape)
    return k.astype(np.float16) / np.float16(255)
  if len(state_shape) == 3:   # low-dim poses, stored as fp16 by the replay memory (Appendix C-7)
    s1 = rs.uniform(-1, 1, (B,) + state_shape).astype(np.float16)
    s2 = rs.uniform(-1, 1, (B,) + state_shape).astype(np.float16)
  else:
    s1, s2 = states(), states()
  a = rs.uniform(-1, 1, (B, 2)).astype(np.float32)
  r = rs.uniform(0, 2, (B, 1)).astype(np.float32)
  m = (rs.rand(B, 1) > 0.25).astype(np.float32)
  return s1, a, r, m, s2


def _np(x):
  return x.detach().numpy() if torch.is_tensor(x) else np.asarray(x)


def ddpg_params(rs, shape, pixels, perturb_targets=True, **kw):
  P = {}
  bn = bool(kw.get("batch_norm", False))
  a = no.ddpg_actor("actor", shape, pixels, kw.get("actor_hidden", "100,100,50"), batch_norm=bn)
  c = no.ddpg_critic("critic", shape, pixels, kw.get("critic_hidden", "100,100,50"), batch_norm=bn)
  for d in (a, c):
    P.update(no.init_params(d, rs))
  # realistic non-zero biases / bigger action head so every gradient path is exercised
  for k in list(P):
    if k.endswith("/moving_mean"):       # the reference never moves them off 0 / 1; off-default values pin that inference reads them
      P[k] = torch.tensor(rs.uniform(-0.05, 0.05, tuple(P[k].shape)).astype(np.float32), dtype=torch.float64)
    if k.endswith("/moving_variance"):
      P[k] = torch.tensor(rs.uniform(0.8, 1.2, tuple(P[k].shape)).astype(np.float32), dtype=torch.float64)
    if k.endswith("biases") or k.endswith("/beta"):
      P[k] = torch.tensor(rs.uniform(-0.1, 0.1, tuple(P[k].shape)).astype(np.float32), dtype=torch.float64)
    if k == "actor/output_action/weights":
      P[k] = torch.tensor(rs.uniform(-0.3, 0.3, tuple(P[k].shape)).astype(np.float32), dtype=torch.float64)
  for src, dst in (("actor", "target_actor"), ("critic", "target_critic")):
    T = no.retarget({k: v for k, v in P.items() if k.startswith(src + "/")}, src, dst)
    if perturb_targets:
      for k in T:
        T[k] = torch.tensor((T[k].numpy() + rs.uniform(-0.02, 0.02, tuple(T[k].shape))).astype(np.float32), dtype=torch.float64)
    P.update(T)
  return P


def golden_ddpg(name, shape, pixels, B, seed, sparse=False, batch_norm=False):
  rs = np.random.RandomState(seed)
  P = ddpg_params(rs, shape, pixels, batch_norm=batch_norm)
  out = {"meta": json.dumps(dict(state_shape=shape, pixels=pixels, B=B, seed=seed, sparse=sparse, batch_norm=bool(batch_norm)))}
  for k, v in P.items():
    out["P0/" + k] = v.numpy().astype(np.float32)
  o = no.DDPGOracle(shape, pixels, P, batch_norm=batch_norm)
  for step in range(2):
    batch = _batch(rs, B, shape, sparse)
    for f, v in zip(("s1", "a", "r", "m", "s2"), batch):
      out["step%d/%s" % (step, f)] = v
    loss0, td0, q0 = o.check_loss(batch)
    out["step%d/check_loss" % step] = _np(loss0); out["step%d/check_td" % step] = _np(td0); out["step%d/check_q" % step] = _np(q0)
    ra = o.actor_train(batch[0])
    out["step%d/actor_grads" % step] = _np(torch.cat([g.reshape(-1) for g in ra["grads"]]))
    out["step%d/actor_norm" % step] = _np(ra["norm"]); out["step%d/mu" % step] = _np(ra["mu"]); out["step%d/dqda" % step] = _np(ra["dqda"])
    rc = o.critic_train(batch)
    out["step%d/critic_grads" % step] = _np(torch.cat([g.reshape(-1) for g in rc["grads"]]))
    out["step%d/critic_norm" % step] = _np(rc["norm"]); out["step%d/loss" % step] = _np(rc["loss"])
    out["step%d/td" % step] = _np(rc["td"]); out["step%d/q" % step] = _np(rc["q"])
    o.update_targets(0.05)
  for k, v in o.P.items():
    out["Pfinal/%s" % k] = v.detach().numpy()
  out["action_given0"] = _np(o.action_given(batch[0][0]))
  np.savez_compressed(os.path.join(GOLD, "nets_%s.npz" % name), **out)
  print("nets_%s" % name)


def golden_naf(name, shape, pixels, B, seed, optimiser, optimiser_args, share=False, batch_norm=False):
  rs = np.random.RandomState(seed)
  P = {}
  value = no.naf_value("value", shape, pixels, batch_norm=batch_norm)
  heads = no.naf_shared_heads(value.fc[-2].out) if share else (no.naf_mu(shape, pixels, batch_norm=batch_norm),
                                                               no.naf_l(shape, pixels, batch_norm=batch_norm))
  defs = [value] + list(heads)
  for d in defs:
    P.update(no.init_params(d, rs))
  for k in list(P):
    if k.endswith("/moving_mean"):
      P[k] = torch.tensor(rs.uniform(-0.05, 0.05, tuple(P[k].shape)).astype(np.float32), dtype=torch.float64)
    if k.endswith("/moving_variance"):
      P[k] = torch.tensor(rs.uniform(0.8, 1.2, tuple(P[k].shape)).astype(np.float32), dtype=torch.float64)
    if k.endswith("biases") or k.endswith("/beta"):
      P[k] = torch.tensor(rs.uniform(-0.1, 0.1, tuple(P[k].shape)).astype(np.float32), dtype=torch.float64)
    if k == "naf/output_action/fc/weights":
      P[k] = torch.tensor(rs.uniform(-0.3, 0.3, tuple(P[k].shape)).astype(np.float32), dtype=torch.float64)
  T = no.retarget({k: v for k, v in P.items() if k.startswith("value/")}, "value", "target_value")
  for k in T:
    T[k] = torch.tensor((T[k].numpy() + rs.uniform(-0.02, 0.02, tuple(T[k].shape))).astype(np.float32), dtype=torch.float64)
  P.update(T)
  out = {"meta": json.dumps(dict(state_shape=shape, pixels=pixels, B=B, seed=seed, optimiser=optimiser, optimiser_args=optimiser_args,
                                 share=bool(share), batch_norm=bool(batch_norm)))}
  for k, v in P.items():
    out["P0/" + k] = v.numpy().astype(np.float32)
  o = no.NAFOracle(shape, pixels, P, optimiser=optimiser, optimiser_args=optimiser_args, share=share, batch_norm=batch_norm)
  for step in range(3):
    batch = _batch(rs, B, shape)
    for f, v in zip(("s1", "a", "r", "m", "s2"), batch):
      out["step%d/%s" % (step, f)] = v
    dv = o.debug_values(batch)
    for f, v in zip(("l_values", "dbg_loss", "V", "A", "V2"), dv):
      out["step%d/%s" % (step, f)] = v
    r = o.train(batch)
    out["step%d/loss" % step] = _np(r["loss"]); out["step%d/norm" % step] = _np(r["norm"])
    out["step%d/grads" % step] = _np(torch.cat([g.reshape(-1) for g in r["grads"]]))
    o.update_targets(0.05)
  for k, v in o.P.items():
    out["Pfinal/%s" % k] = v.detach().numpy()
  out["action_given0"] = _np(o.action_given(batch[0][0]))
  np.savez_compressed(os.path.join(GOLD, "nets_%s.npz" % name), **out)
  print("nets_%s" % name)


def golden_lrpg(seed=31):
  rs = np.random.RandomState(seed)
  shape = (2, 2, 7)
  P = no.init_params(no.lrpg_model(shape), rs)
  for k in list(P):
    if k.endswith("biases"):
      P[k] = torch.tensor(rs.uniform(-0.1, 0.1, tuple(P[k].shape)).astype(np.float32), dtype=torch.float64)
  out = {"meta": json.dumps(dict(state_shape=shape, seed=seed))}
  for k, v in P.items():
    out["P0/" + k] = v.numpy().astype(np.float32)
  o = no.LRPGOracle(shape, P, optimiser="Adam", optimiser_args={"learning_rate": 0.01})
  for step, N in enumerate((37, 1, 203)):
    obs = rs.uniform(-1, 1, (N,) + shape).astype(np.float32)
    act = rs.randint(0, 5, N).astype(np.int32)
    adv = rs.uniform(5, 200, N).astype(np.float32) if N > 1 else np.array([3.0], np.float32)
    if N == 1:
      # standardise() divides by a zero std for one sample -> NaNs; the reference skips training when all
      # totals are equal (lrpg_cartpole.py:213-216); keep the case out of the golden step sequence
      continue
    out["step%d/obs" % step] = obs; out["step%d/act" % step] = act; out["step%d/adv" % step] = adv
    r = o.train(obs, act, adv)
    out["step%d/loss" % step] = _np(r["loss"]); out["step%d/norm" % step] = _np(r["norm"])
    out["step%d/logits" % step] = _np(r["logits"])
    out["step%d/grads" % step] = _np(torch.cat([g.reshape(-1) for g in r["grads"]]))
  for k, v in o.P.items():
    out["Pfinal/%s" % k] = v.detach().numpy()
  np.savez_compressed(os.path.join(GOLD, "nets_lrpg.npz"), **out)
  print("nets_lrpg")


def main():
  os.makedirs(GOLD, exist_ok=True)
  ref = load_reference_replay()
  golden_replay(ref)
  golden_mt()
  golden_ddpg("ddpg_pixel", (16, 16, 3, 1, 2), True, 8, 21)
  golden_ddpg("ddpg_pixel_odd", (22, 18, 3, 2, 1), True, 5, 22, sparse=True)   # odd pooling sizes: 22->11->5->2, 18->9->4->2
  golden_ddpg("ddpg_lowdim", (2, 2, 7), False, 16, 23)
  golden_naf("naf_pixel", (16, 16, 3, 2, 1), True, 8, 24, "Adam", {"learning_rate": 0.01})
  golden_naf("naf_lowdim", (3, 2, 7), False, 16, 25, "Momentum", {"learning_rate": 0.01, "momentum": 0.9})
  golden_naf_shared()
  golden_batch_norm()
  golden_lrpg()


def golden_batch_norm():
  """--use-batch-norm (base_network.py:74-79): oracle-side vectors for the round that builds the kernels"""
  golden_ddpg("ddpg_pixel_bn", (16, 16, 3, 1, 2), True, 8, 28, batch_norm=True)
  # NAF with Momentum + batch norm as in exps/run_84.sh:11-15, here combined with the shared representation so that both graph
  # variants meet (three heads on one batch-normalised trunk)
  golden_naf("naf_pixel_bn_shared", (16, 16, 3, 2, 1), True, 8, 29, "Momentum", {"learning_rate": 0.01, "momentum": 0.9},
             share=True, batch_norm=True)


def golden_naf_shared():
  """--share-input-state-representation (naf_cartpole.py:151-154,176-179)"""
  # (plain SGD here: Adam's g / (|g| + 1e-8) update is ill-conditioned wherever the summed three-head gradient nearly cancels,
  # which says nothing about the shared graph; Adam itself is pinned by naf_pixel)
  golden_naf("naf_pixel_shared", (16, 16, 3, 2, 1), True, 8, 26, "GradientDescent", {"learning_rate": 0.01}, share=True)
  golden_naf("naf_lowdim_shared", (3, 2, 7), False, 16, 27, "Momentum", {"learning_rate": 0.01, "momentum": 0.9}, share=True)


if __name__ == "__main__":
  main()
