"""Host-side mirror of the reference's base_network.py (/root/reference/base_network.py:13-134).

The reference builds a TensorFlow graph; here the same builder calls record a *network description*
(conv trunk + FC stack) that is handed to libcartpolepp (cpp_net_spec, include/cartpolepp.h), which owns
every numeric operation.  Names, argument order and error behaviour follow the reference:

  Network(namespace)                                   base_network.py:16-18
  ._create_variables_copy_op / .set_as_target_network_for / .update_weights      :20-49
  .trainable_model_vars                                :51-56
  .hidden_layers_starting_at(layer, layer_sizes, opts) :58-71
  .simple_conv_net_on(input_layer, opts)               :73-127
  .input_state_network(input_state, opts)              :129-134
"""
import collections
import math
import sys
import ctypes as C
import numpy as np

from . import _lib

# base_network.py:11 - a global flag fed to every Session.run.  It only matters for dropout / batch norm; the library's
# agent entry points set it per call exactly as the reference's feed_dicts do (train ops True, everything else False).
IS_TRAINING = False

ACT = {None: 0, "relu": 1, "tanh": 2}


class Placeholder(object):
  """stands in for tf.placeholder(shape=[None]+state_shape): only the per-sample shape matters"""

  def __init__(self, shape, name=None):
    self.shape = tuple(int(d) for d in shape if d is not None)
    self.name = name

  def get_shape(self):
    return (None,) + self.shape


class Layer(object):
  """symbolic output of a builder call: the layers stacked so far on top of a Placeholder"""

  def __init__(self, source, conv=False, fc=(), flat=False, concat=None, bn=False):
    self.source = source          # Placeholder
    self.bn = bool(bn)            # --use-batch-norm: slim.batch_norm after every conv of the trunk (base_network.py:74-79)
    self.conv = conv              # True once simple_conv_net_on was applied
    self.fc = list(fc)            # [(scope, out, act, dropout)]
    self.flat = flat
    self.concat = concat          # (index of the FC layer the action is concatenated in front of, action_dim)

  def feature_shape(self):
    if self.conv:
      h, w = self.source.shape[0], self.source.shape[1]
      for _ in range(3):
        h, w = h // 2, w // 2
      return (h, w, 10)
    return self.source.shape


def fully_connected(layer, num_outputs, scope, activation="relu", dropout=False):
  """slim.fully_connected on a (flattened) symbolic layer; dropout: followed by slim.dropout(keep_prob 0.5, is_training=IS_TRAINING)"""
  if not isinstance(layer, Layer):
    layer = Layer(layer)
  return Layer(layer.source, layer.conv, layer.fc + [(scope, int(num_outputs), activation, bool(dropout))], True, layer.concat, layer.bn)


def flatten(layer):
  if not isinstance(layer, Layer):
    layer = Layer(layer)
  return Layer(layer.source, layer.conv, layer.fc, True, layer.concat, layer.bn)


def concat_action(layer, action_dim):
  """tf.concat(1, [layer, action]) (ddpg_cartpole.py:170,175)"""
  if not isinstance(layer, Layer):
    layer = Layer(layer)
  assert layer.concat is None
  return Layer(layer.source, layer.conv, layer.fc, True, (len(layer.fc), int(action_dim)), layer.bn)


Var = collections.namedtuple("Var", "name shape offset size")


class Network(object):
  """Common class for handling ops for making / updating target networks."""

  def __init__(self, namespace):
    self.namespace = namespace
    self.target_update_op = None
    self.update_weights_op = None     # Appendix C-5: the reference forgets to initialise this one
    self._final = None                # Layer once the subclass finished building
    self._spec = None
    self._engine = None               # set by the agent-level engine that owns the parameters
    self._part = None                 # name of this network's slice inside the engine's flat buffers

  # ---- graph-builder mirror ------------------------------------------------------------------
  def hidden_layers_starting_at(self, layer, layer_sizes, opts=None):
    if not isinstance(layer_sizes, list):
      layer_sizes = [int(s) for s in layer_sizes.split(",")]
    assert len(layer_sizes) > 0
    use_dropout = opts is not None and bool(getattr(opts, "use_dropout", False))     # Appendix C-1: opts=None means no dropout
    for i, size in enumerate(layer_sizes):
      layer = fully_connected(layer, size, scope="h%d" % i, activation="relu", dropout=use_dropout)   # slim.dropout scope "do%d", :69-70
    return layer

  def simple_conv_net_on(self, input_layer, opts):
    src = input_layer.source if isinstance(input_layer, Layer) else input_layer
    if len(src.shape) < 3:
      raise ValueError("simple_conv_net_on needs a (H, W, ...) state, got %s" % (src.shape,))
    height, width = src.shape[0], src.shape[1]
    num_channels = int(np.prod(src.shape[2:]))
    sys.stderr.write("input_layer (?, %d, %d, %d) #%d\n" % (height, width, num_channels, height * width * num_channels))
    # normalizer_fn=slim.batch_norm, normalizer_params={'is_training': IS_TRAINING} (base_network.py:74-79)
    return Layer(src, conv=True, bn=bool(getattr(opts, "use_batch_norm", False)))

  def input_state_network(self, input_state, opts):
    if opts.use_raw_pixels:
      input_state = self.simple_conv_net_on(input_state, opts)
    flattened_input_state = flatten(input_state)
    return self.hidden_layers_starting_at(flattened_input_state, opts.hidden_layers, opts)

  # ---- description -> cpp_net_spec ---------------------------------------------------------------
  def _finalise(self, layer):
    self._final = layer
    src = layer.source
    spec = _lib.NetSpec()
    spec.pixels = 1 if layer.conv else 0
    if layer.conv:
      spec.H, spec.W, spec.Cin = src.shape[0], src.shape[1], int(np.prod(src.shape[2:]))
      spec.input_dim = 0
    else:
      spec.input_dim = int(np.prod(src.shape))
    if len(layer.fc) > _lib.CPP_MAX_FC:
      raise ValueError("at most %d fully connected layers" % _lib.CPP_MAX_FC)
    spec.n_fc = len(layer.fc)
    for i, (_, out, act, drop) in enumerate(layer.fc):
      spec.fc_out[i], spec.fc_act[i], spec.fc_dropout[i] = out, ACT[act], int(drop)
    spec.concat_at = layer.concat[0] if layer.concat else -1
    spec.action_dim = layer.concat[1] if layer.concat else 0
    spec.batch_norm = 1 if (layer.conv and layer.bn) else 0
    self._spec = spec
    return spec

  def _variables(self):
    """[Var] in TF creation order; offsets are relative to this network's slice"""
    layer, out, off = self._final, [], 0
    def add(name, shape):
      nonlocal off
      n = int(np.prod(shape))
      out.append(Var("%s/%s" % (self.namespace, name), tuple(shape), off, n))
      off += n
    d = None
    if layer.conv:
      cin = int(np.prod(layer.source.shape[2:]))
      for name, k in (("conv1", 5), ("conv2", 5), ("conv3", 3)):
        add(name + "/weights", (k, k, cin, 10))
        if layer.bn:     # slim.conv2d creates no bias under a normalizer_fn; slim.batch_norm(center=True, scale=False)
          add(name + "/BatchNorm/beta", (10,)); add(name + "/BatchNorm/moving_mean", (10,)); add(name + "/BatchNorm/moving_variance", (10,))
        else:
          add(name + "/biases", (10,))
        cin = 10
    d = int(np.prod(layer.feature_shape()))
    for i, (scope, o, _, _) in enumerate(layer.fc):
      if layer.concat and layer.concat[0] == i:
        d += layer.concat[1]
      add(scope + "/weights", (d, o)); add(scope + "/biases", (o,))
      d = o
    return out

  def num_params(self):
    return sum(v.size for v in self._variables())

  def initial_values(self, rng, small_uniform=()):
    """xavier-uniform weights, zero biases (slim defaults); scopes in `small_uniform` get U(+-1e-3)
    (ddpg_cartpole.py:94, naf_cartpole.py:155).  Returns a flat float32 vector."""
    flat = np.zeros(self.num_params(), dtype=np.float32)
    for v in self._variables():
      if v.name.endswith("/moving_variance"):
        flat[v.offset:v.offset + v.size] = 1.0      # slim.batch_norm initialiser; never updated by the reference (Appendix A-5)
        continue
      if v.name.endswith("/biases") or v.name.endswith("/beta") or v.name.endswith("/moving_mean"):
        continue
      if len(v.shape) == 4:
        kh, kw, ci, co = v.shape
        lim = math.sqrt(6.0 / (kh * kw * ci + kh * kw * co))
      else:
        lim = math.sqrt(6.0 / (v.shape[0] + v.shape[1]))
      if any(v.name.endswith("/%s/weights" % s) for s in small_uniform):
        lim = 1e-3
      flat[v.offset:v.offset + v.size] = rng.uniform(-lim, lim, v.size).astype(np.float32)
    return flat

  # ---- parameters live in the engine's flat device buffer ------------------------------------------
  def _need_engine(self):
    if self._engine is None:
      raise RuntimeError("network '%s' is not attached to a training engine yet" % self.namespace)
    return self._engine

  def flat_params(self):
    """torch view (device) of this network's parameters"""
    return self._need_engine().part_view(self._part)

  def get_variable(self, name):
    """torch view of one variable, by its reference name e.g. 'actor/conv1/weights'"""
    for v in self._variables():
      if v.name == name or v.name == "%s/%s" % (self.namespace, name):
        return self.flat_params()[v.offset:v.offset + v.size].view(v.shape)
    raise KeyError(name)

  def set_variables(self, values):
    """values: {name: array}; copies into the device buffer (used to inject weights for parity tests)"""
    import torch
    flat = self.flat_params()
    for v in self._variables():
      if v.name in values:
        a = np.ascontiguousarray(np.asarray(values[v.name], dtype=np.float32).reshape(-1))
        assert a.size == v.size, (v.name, a.size, v.size)
        flat[v.offset:v.offset + v.size].copy_(torch.from_numpy(a))

  def trainable_model_vars(self):
    return [v for v in self._variables() if v.name.startswith(self.namespace)]

  # ---- target network handling -------------------------------------------------------------------
  def _create_variables_copy_op(self, source_network, affine_combo_coeff):
    assert affine_combo_coeff >= 0.0 and affine_combo_coeff <= 1.0
    src_vars, dst_vars = source_network._variables(), self._variables()
    assert [v.shape for v in src_vars] == [v.shape for v in dst_vars]
    return (source_network, float(affine_combo_coeff))

  def _run_copy_op(self, op):
    source_network, coeff = op
    import torch
    t, s = self.flat_params(), source_network.flat_params()
    _lib.check(_lib.lib().cpp_soft_update(_lib.ptr(t), _lib.ptr(s), C.c_float(coeff), C.c_int64(t.numel()),
                                          _lib.stream_ptr()))

  def set_as_target_network_for(self, source_network, target_update_rate):
    """Create an op that will update this networks weights based on a source_network"""
    # one off: theta' <- theta' - 1.0*(theta' - theta), exactly the reference's assign_sub (Appendix A-11)
    self._run_copy_op(self._create_variables_copy_op(source_network, 1.0))
    self.update_weights_op = self._create_variables_copy_op(source_network, target_update_rate)

  def update_weights(self):
    """called during training to update target network."""
    if self.update_weights_op is None:
      raise Exception("not a target network? or set_source_network not yet called")
    return self._run_copy_op(self.update_weights_op)
