"""__graft_entry__.smoke(): one small DDPG pixel grad-step on cuda:0 through the drop-in classes, checked against the fp64
oracle: forward observables, EVERY gradient tensor (oracle evaluated with the GPU's own max-pool / ReLU routing, so a wrong
gradient cannot hide behind lr * g) and the parameters after the step."""
import json
import numpy as np
import torch


def run_smoke():
  assert torch.cuda.is_available(), "smoke() needs a GPU"
  from tests import gpu_util as U
  from oracle import nets_oracle as no
  from oracle.make_golden import ddpg_params, _batch
  parts = [(0, "actor"), (1, "critic")]
  shape, B = (32, 32, 3, 1, 2), 16
  rs = np.random.RandomState(1)
  P = ddpg_params(rs, shape, True)
  batch = _batch(rs, B, shape)
  nets, eng, o = U.make_ddpg(shape, True, {k: v.numpy() for k, v in P.items()}, batch_size=B)
  orc = no.DDPGOracle(shape, True, P)
  b = U.Batch(*batch)
  err = {}
  l0, td0, q0 = orc.check_loss(batch)
  loss, td, q = nets["critic"].check_loss(b)
  err["check_loss"] = max(U.assert_close(loss, l0.numpy(), what="loss"), U.assert_close(td, td0.numpy(), what="td"),
                          U.assert_close(q, q0.numpy(), what="q"))
  eng.actor_backward(b.state_1)
  with no.gates(U.conv_routing(eng, parts, shape, B)) as stats:
    ra = orc.actor_train(batch[0])
  U.check_gate_stats(stats, {})
  rep = U.per_variable_errors(U.names_of(nets["actor"]), eng.buffers["grads"][:eng.n_actor].cpu().numpy(), [x.numpy() for x in ra["grads"]])
  eng.actor_apply()
  eng.critic_backward(b, reuse_s1_trunk=True)
  with no.gates(U.conv_routing(eng, parts[1:], shape, B)) as stats:
    rc = orc.critic_train(batch)
  U.check_gate_stats(stats, {})
  rep.update(U.per_variable_errors(U.names_of(nets["critic"]), eng.buffers["grads"][eng.off_critic:eng.off_critic + eng.n_critic].cpu().numpy(),
                                   [x.numpy() for x in rc["grads"]]))
  err["worst_gradient_tensor"] = U.assert_all_within(rep, "smoke: DDPG gradients")
  err["loss"] = U.assert_close(eng.last_loss(), float(rc["loss"]), what="loss")
  eng.critic_apply()
  for k in ("actor", "critic"):
    want = np.concatenate([orc.P[n].numpy().reshape(-1) for n in U.names_of(nets[k])])
    err["P_" + k] = U.assert_close(U.flat_of(nets[k]), want, what="params " + k)
  torch.cuda.synchronize()
  print("smoke ok: DDPG pixel grad-step on %s, rel err vs fp64 oracle %s" % (torch.cuda.get_device_name(0), json.dumps(err)))
