"""Mirror of the reference's util.py for the pieces the hot path and its callers need
(/root/reference/util.py): flags :10-20, construct_optimiser :73-76 (returns the optimiser *description* the CUDA step consumes),
OrnsteinUhlenbeckNoise :134-156 (host side, including the rotated np.clip arguments, Appendix C-6),
SaverUtil :88-131 (checkpoint save / restore of the device-resident variables)."""
import datetime
import json
import os
import sys
import time
import numpy as np

from ._engine import parse_optimiser


def add_opts(parser):
  parser.add_argument('--gradient-clip', type=float, default=5, help="do global clipping to this norm")
  parser.add_argument('--print-gradients', action='store_true', help="whether to verbose print all gradients and l2 norms")
  parser.add_argument('--optimiser', type=str, default="GradientDescent", help="tf.train.XXXOptimizer to use")
  parser.add_argument('--optimiser-args', type=str, default="{\"learning_rate\": 0.001}",
                      help="json serialised args for optimiser constructor")
  parser.add_argument('--use-dropout', action='store_true', help="include a dropout layers after each fully connected layer")


def construct_optimiser(opts):
  """-> (kind, hyper-parameter dict) for libcartpolepp's optimiser kernels"""
  return parse_optimiser(opts.optimiser, json.loads(opts.optimiser_args))


def shape_and_product_of(shape):
  n = 1
  for d in shape:
    if d is not None:
      n *= int(d)
  return "%s #%s" % (tuple(shape), n)


class SaverUtil(object):
  """util.py:88-131.  The reference hands `tf.train.Saver` every variable except the replay memory; here the variables
  are the engine's flat device buffers: each network's variables go into one `.npz` under their reference names
  (`actor/conv1/weights` ...), optimiser slots and the Adam power accumulators under `__slots__` / `__opt_state__`.
  Like the Saver, a text file `checkpoint` in the directory names the latest save (`model_checkpoint_path: "ckpt.<ts>"`);
  restore copies into the bound buffers in place, so captured CUDA graphs stay valid."""

  def __init__(self, engine, ckpt_dir="/tmp", save_freq=60):
    self.engine = engine
    self.ckpt_dir = ckpt_dir
    if not os.path.exists(self.ckpt_dir):
      os.makedirs(self.ckpt_dir)
    assert save_freq > 0
    self.save_freq = save_freq
    self.load_latest_ckpt_or_init_if_none()

  def _networks(self):
    e = self.engine
    nets = getattr(e, "nets", None)
    return list(nets.items()) if nets is not None else [("model", e.agent)]

  def _latest(self):
    ckpt_info_file = "%s/checkpoint" % self.ckpt_dir
    if not os.path.isfile(ckpt_info_file):
      return None
    for line in open(ckpt_info_file, "r"):
      key, _, value = line.partition(":")
      if key.strip() == "model_checkpoint_path":
        return value.strip().strip('"')
    raise AssertionError("no model_checkpoint_path in %s" % ckpt_info_file)

  def load_latest_ckpt_or_init_if_none(self):
    """loads latest ckpt from dir; if there is none the (already initialised) variables are saved straight away.
    Data parallel: rank 0 alone reads the directory and restores (or keeps its initialisation); its variables, targets and
    optimiser state are then broadcast, so every replica starts from the same bits whether or not the file system is shared."""
    if self._is_writer():
      latest = self._latest()
      if latest is None:
        sys.stderr.write("no latest ckpt in %s, just initing vars...\n" % self.ckpt_dir)
        self.force_save()
      else:
        most_recent_ckpt = "%s/%s" % (self.ckpt_dir, latest)
        sys.stderr.write("loading ckpt %s\n" % most_recent_ckpt)
        self._restore_local(most_recent_ckpt)
    self._sync_replicas()
    self.next_scheduled_save_time = time.time() + self.save_freq

  def _sync_replicas(self):
    """rank 0's buffers -> every rank (a no-op without an initialised process group), then a barrier"""
    try:
      import torch.distributed as dist
    except ImportError:
      return
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
      return
    for name in ("params", "target_params", "slots", "opt_state"):
      if name in self.engine.buffers:
        dist.broadcast(self.engine.buffers[name], src=0)
    dist.barrier()

  def restore(self, path):
    """restore `path` into the bound buffers in place; data parallel: rank 0 reads, everyone receives its bits"""
    if self._is_writer():
      self._restore_local(path)
    self._sync_replicas()

  def _restore_local(self, path):
    import torch
    e = self.engine
    with np.load(path if path.endswith(".npz") else path + ".npz") as z:
      for part, net in self._networks():
        values = {}
        for v in net._variables():
          if v.name not in z.files:
            raise KeyError("checkpoint %s has no variable %s" % (path, v.name))
          if tuple(z[v.name].shape) != tuple(v.shape):
            raise ValueError("checkpoint %s: %s has shape %s, the model wants %s" % (path, v.name, z[v.name].shape, v.shape))
          values[v.name] = z[v.name]
        net.set_variables(values)
      for key, buf in (("__slots__", "slots"), ("__opt_state__", "opt_state")):
        if buf in e.buffers:
          if key not in z.files or z[key].size != e.buffers[buf].numel():
            raise ValueError("checkpoint %s does not hold this optimiser's %s" % (path, buf))
          e.buffers[buf].copy_(torch.from_numpy(np.ascontiguousarray(z[key], dtype=np.float32)))
    if e.buffers["params"].is_cuda:
      torch.cuda.current_stream().synchronize()

  @staticmethod
  def _is_writer():
    """data parallel: replicas are bit-identical (set_data_parallel / _sync_replicas broadcast rank 0's bits, only summed
    gradients are exchanged afterwards), so rank 0 alone writes.  A process is a non-writer only inside an initialised
    process group, or when the launcher's environment says so (WORLD_SIZE > 1 and RANK != 0): a stray RANK variable in the
    environment of a single-process run must not silence its checkpoints."""
    try:
      import torch.distributed as dist
      if dist.is_available() and dist.is_initialized():
        return dist.get_rank() == 0
    except ImportError:
      pass
    return not (int(os.environ.get("WORLD_SIZE", "1")) > 1 and int(os.environ.get("RANK", "0")) != 0)

  def force_save(self):
    """force a save now."""
    if not self._is_writer():
      sys.stderr.write("rank != 0: checkpoint left to rank 0\n")
      self.next_scheduled_save_time = time.time() + self.save_freq
      return
    dts = datetime.datetime.now().strftime('%Y%m%d_%H%M%S')
    name = "ckpt.%s" % dts
    n = 1
    while os.path.exists("%s/%s.npz" % (self.ckpt_dir, name)):      # two saves within one second
      name = "ckpt.%s_%d" % (dts, n)
      n += 1
    new_ckpt = "%s/%s" % (self.ckpt_dir, name)
    sys.stderr.write("saving ckpt %s\n" % new_ckpt)
    start_time = time.time()
    e = self.engine
    arrays = {}
    for part, net in self._networks():
      flat = e.part_view(part).detach().cpu().numpy()
      for v in net._variables():
        arrays[v.name] = flat[v.offset:v.offset + v.size].reshape(v.shape)
    for key, buf in (("__slots__", "slots"), ("__opt_state__", "opt_state")):
      if buf in e.buffers:
        arrays[key] = e.buffers[buf].detach().cpu().numpy()
    tmp = "%s/.%s.tmp.npz" % (self.ckpt_dir, name)
    np.savez(tmp, **arrays)
    os.replace(tmp, new_ckpt + ".npz")
    history = []
    info = "%s/checkpoint" % self.ckpt_dir
    if os.path.isfile(info):
      history = [l for l in open(info) if l.startswith("all_model_checkpoint_paths")]
    with open(info + ".tmp", "w") as f:
      f.write('model_checkpoint_path: "%s"\n' % name)
      f.writelines(history)
      f.write('all_model_checkpoint_paths: "%s"\n' % name)
    os.replace(info + ".tmp", info)
    print("save_took", time.time() - start_time)
    self.next_scheduled_save_time = time.time() + self.save_freq

  def save_if_required(self):
    """check if save is required based on time and if so, save."""
    if time.time() >= self.next_scheduled_save_time:
      self.force_save()


class OrnsteinUhlenbeckNoise(object):
  """generate time correlated noise for action exploration"""

  def __init__(self, dim, theta=0.01, sigma=0.2, max_magnitude=1.5):
    self.dim, self.theta, self.sigma, self.max_magnitude = dim, theta, sigma, max_magnitude
    self.state = np.zeros(self.dim)

  def sample(self):
    self.state += self.theta * -self.state
    self.state += self.sigma * np.random.randn(self.dim)
    # util.py:155 passes np.clip(max, -max, state): the arguments are rotated in the reference, which
    # evaluates to minimum(state, max) for these values; reproduced as is
    self.state = np.clip(self.max_magnitude, -self.max_magnitude, self.state)
    return np.copy(self.state)
