"""Agent-level engines: own the flat device buffers (torch tensors as containers) and drive the step
functions of libcartpolepp.  One engine per agent; the Network objects of base_network.py are views."""
import ctypes as C
import numpy as np
import torch

from . import _lib

OPTIMISERS = {"GradientDescent": 0, "Momentum": 1, "Adam": 2}


def _require_cuda():
  if not torch.cuda.is_available():
    raise RuntimeError("cartpoleplusplus_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def parse_optimiser(name, args):
  """util.construct_optimiser (util.py:73-76): tf.train.<name>Optimizer(**args) -> (kind, hyper-parameters)"""
  if name not in OPTIMISERS:
    raise ValueError("tf.train.%sOptimizer is not implemented on this path (have %s)" % (name, sorted(OPTIMISERS)))
  a = dict(args)
  hp = dict(lr=float(a.pop("learning_rate")), momentum=0.0, beta1=0.9, beta2=0.999, eps=1e-8)
  if name == "Momentum":
    hp["momentum"] = float(a.pop("momentum"))
  if name == "Adam":
    hp["beta1"] = float(a.pop("beta1", 0.9)); hp["beta2"] = float(a.pop("beta2", 0.999)); hp["eps"] = float(a.pop("epsilon", 1e-8))
  a.pop("use_locking", None); a.pop("name", None)
  if a:
    raise TypeError("unexpected optimiser args %s" % sorted(a))
  return OPTIMISERS[name], hp


class DeviceStager(object):
  """host -> device staging of step inputs through reusable pinned buffers (async on the current stream)"""

  def __init__(self, device):
    self.device = device
    self._pinned = {}
    self._dev = {}
    self._events = {}

  def __call__(self, key, x, dtype=None):
    """returns a device tensor holding x (numpy / torch cpu / torch cuda); dtype None keeps fp16/fp32 as is"""
    if torch.is_tensor(x):
      if x.is_cuda:
        t = x if dtype is None or x.dtype == dtype else x.to(dtype)
        return t.contiguous()
      src = x
    else:
      a = np.asarray(x)
      if a.dtype not in (np.float16, np.float32, np.int32, np.int64):
        a = a.astype(np.float32)
      src = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and src.dtype != dtype:
      src = src.to(dtype)
    src = src.contiguous()
    n, k = src.numel(), (key, src.dtype)
    dev = self._dev.get(k)
    if dev is None or dev.numel() < n:
      dev = torch.empty(max(n, 1), dtype=src.dtype, device=self.device)
      self._dev[k] = dev
    out = dev[:n].view(src.shape)
    if src.is_pinned():
      out.copy_(src, non_blocking=True)
    else:
      pin = self._pinned.get(k)
      if pin is None or pin.numel() < n:
        pin = torch.empty(max(n, 1), dtype=src.dtype, pin_memory=True)
        self._pinned[k] = pin
      # the pinned buffer is reused: wait until the previous async copy out of it has finished
      ev = self._events.get(k)
      if ev is None:
        ev = self._events[k] = torch.cuda.Event()
      else:
        ev.synchronize()
      pin[:n].view(src.shape).copy_(src)
      out.copy_(pin[:n].view(src.shape), non_blocking=True)
      ev.record()
    return out


def state_flag(t):
  if t.dtype == torch.float16:
    return 1
  if t.dtype == torch.float32:
    return 0
  raise TypeError("states must be float16 (replay memory) or float32 (environment), got %s" % t.dtype)


class EngineBase(object):
  def __init__(self):
    _require_cuda()
    self.lib = _lib.lib()
    self.device = torch.device("cuda", torch.cuda.current_device())
    self.stage = DeviceStager(self.device)
    self.parts = {}           # part name -> (buffer name, offset, size)
    self.buffers = {}

  def part_view(self, part):
    bufname, off, n = self.parts[part]
    return self.buffers[bufname][off:off + n]

  def _stream(self):
    return _lib.stream_ptr()

  @staticmethod
  def _f(x):
    return C.c_float(float(x))
