"""__graft_entry__.smoke(): one small DDPG pixel grad-step on cuda:0 through the drop-in classes,
checked against the fp64 oracle."""
import numpy as np
import torch


def run_smoke():
  assert torch.cuda.is_available(), "smoke() needs a GPU"
  from tests import gpu_util as U
  from oracle import nets_oracle as no
  from oracle.make_golden import ddpg_params, _batch
  shape, B = (32, 32, 3, 1, 2), 16
  rs = np.random.RandomState(1)
  P = ddpg_params(rs, shape, True)
  batch = _batch(rs, B, shape)
  nets, eng, o = U.make_ddpg(shape, True, {k: v.numpy() for k, v in P.items()}, batch_size=B)
  orc = no.DDPGOracle(shape, True, P)
  b = U.Batch(*batch)
  nets["actor"].train(b.state_1)
  nets["critic"].train(b)
  orc.actor_train(batch[0]); rc = orc.critic_train(batch)
  err = {}
  for k in ("actor", "critic"):
    want = np.concatenate([orc.P[n].numpy().reshape(-1) for n in U.names_of(nets[k])])
    err[k] = U.assert_close(U.flat_of(nets[k]), want, what="params " + k)
  err["loss"] = U.assert_close(eng.last_loss(), float(rc["loss"]), what="loss")
  torch.cuda.synchronize()
  print("smoke ok: DDPG pixel grad-step on %s, rel err vs fp64 oracle %s" % (torch.cuda.get_device_name(0), err))
