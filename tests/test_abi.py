"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/cartpolepp.h
declares, and its host-only entry points (index sampling, network layout) agree with the oracle.
No GPU compute is invoked here."""
import ctypes as C
import json
import os
import re
import numpy as np
import pytest

from cartpoleplusplus_b200 import _lib
from oracle import nets_oracle as no

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
  src = open(os.path.join(ROOT, "include", "cartpolepp.h")).read()
  src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
  return sorted(set(re.findall(r"\b(cpp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
  lib = _lib.lib()
  declared = header_symbols()
  assert len(declared) > 50
  for name in declared:
    assert hasattr(lib, name), "libcartpolepp.so does not export %s" % name
  assert sorted(_lib.SYMBOLS) == declared, "python binding list and header disagree"
  assert lib.cpp_version() == 4


def test_error_reporting_is_c_abi_clean():
  lib = _lib.lib()
  out = np.zeros(4, dtype=np.int64)
  h = C.c_void_p()
  _lib.check(lib.cpp_mt_create(C.byref(h)))
  st = lib.cpp_mt_randint(h, C.c_int64(0), C.c_int64(4), out.ctypes.data_as(C.c_void_p))     # numpy raises ValueError(low >= high)
  assert st == -1 and b"high" in lib.cpp_last_error()
  with pytest.raises(_lib.CppError):
    _lib.check(st)
  lib.cpp_mt_destroy(h)


def test_runtime_switches_are_known_and_unknown_names_are_rejected():
  """cpp_set_option (include/cartpolepp.h): every documented switch is accepted without a GPU (they only select routes for later
  launches), defaults are restored, an unknown name is an error with a message - not a silent no-op"""
  lib = _lib.lib()
  header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "cartpolepp.h")).read()
  defaults = {"conv1_tc": -1, "fused_mlp": -1, "prep_hoist": 1, "conv1_split": 1, "critic_tail": 1, "bwd_critic_sms": 92, "fwd_actor_sms": 37,
              "wgrad_tc": 5, "conv_row": 1, "mlp_fast": 3, "wgrad_flush_steps": 32, "fc_tc": 0, "is_training": 1, "dropout_external": 0}
  for name, value in defaults.items():
    assert '"%s"' % name in header, "switch %s is not documented in the header" % name
    _lib.check(lib.cpp_set_option(name.encode(), value))
  st = lib.cpp_set_option(b"no_such_switch", 1)
  assert st != 0 and b"no_such_switch" in lib.cpp_last_error()


def _randint(lib, h, high, n):
  out = np.empty(n, dtype=np.int64)
  _lib.check(lib.cpp_mt_randint(h, C.c_int64(high), C.c_int64(n), out.ctypes.data_as(C.c_void_p)))
  return out


def test_index_sampling_bit_exact_vs_numpy_kat(golden_dir):
  """a1: bit-exact integer parity with np.random.randint (the reference's call, replay_memory.py:123-129)"""
  lib = _lib.lib()
  k = json.load(open(os.path.join(golden_dir, "mt19937_kat.json")))
  h = C.c_void_p()
  _lib.check(lib.cpp_mt_create(C.byref(h)))
  for case in k["kat"]:
    _lib.check(lib.cpp_mt_seed(h, C.c_uint32(case["seed"])))
    for want in case["outs"]:
      assert _randint(lib, h, case["high"], case["n"]).tolist() == want, case["seed"]
  # live against numpy at the reference's default size, and across the 624-word twist boundary
  for seed in (0, 1, 2 ** 32 - 1):
    _lib.check(lib.cpp_mt_seed(h, C.c_uint32(seed)))
    rs = np.random.RandomState(seed)
    for n in (256, 1, 1024, 5000):
      assert np.array_equal(_randint(lib, h, 22000, n), rs.randint(0, 22000, n))
  lib.cpp_mt_destroy(h)


def test_index_sampling_shares_numpys_global_stream(golden_dir):
  """interleaved with other legacy draws (OU noise randn, epsilon-greedy random) exactly like the reference loop"""
  lib = _lib.lib()
  k = json.load(open(os.path.join(golden_dir, "mt19937_kat.json")))["interleaved"]
  h = C.c_void_p()
  _lib.check(lib.cpp_mt_create(C.byref(h)))
  np.random.seed(k["seed"])
  for kind, want in k["seq"]:
    if kind == "randn":
      assert np.random.randn(2).tolist() == want
    elif kind == "random":
      assert float(np.random.random_sample()) == want
    else:
      st = np.random.get_state()
      key = np.ascontiguousarray(st[1], dtype=np.uint32)
      _lib.check(lib.cpp_mt_set_state(h, key.ctypes.data_as(C.c_void_p), C.c_int32(int(st[2]))))
      got = _randint(lib, h, 22000, len(want))
      pos = C.c_int32()
      _lib.check(lib.cpp_mt_get_state(h, key.ctypes.data_as(C.c_void_p), C.byref(pos)))
      np.random.set_state((st[0], key, int(pos.value), st[3], st[4]))
      assert got.tolist() == want
  lib.cpp_mt_destroy(h)


def _spec_of(nd):
  s = _lib.NetSpec()
  s.pixels = 1 if nd.pixels else 0
  if nd.pixels:
    s.H, s.W, s.Cin = nd.H, nd.W, nd.cin
  else:
    s.input_dim = nd.feat
  s.n_fc = len(nd.fc)
  for i, l in enumerate(nd.fc):
    s.fc_out[i] = l.out
    s.fc_act[i] = {None: 0, "relu": 1, "tanh": 2}[l.act]
  s.concat_at = -1 if nd.concat_at is None else nd.concat_at
  s.action_dim = nd.action_dim
  return s


@pytest.mark.parametrize("nd", [
    no.ddpg_actor("actor", (64, 64, 3, 1, 3), True), no.ddpg_critic("critic", (64, 64, 3, 1, 3), True),
    no.ddpg_actor("actor", (2, 2, 7), False), no.ddpg_critic("critic", (2, 2, 7), False),
    no.naf_value("value", (64, 64, 3, 2, 3), True), no.naf_l((128, 128, 3, 2, 4), True),
    no.ddpg_critic("critic", (128, 128, 3, 2, 4), True), no.lrpg_model((2, 2, 7)),
    no.ddpg_actor("actor", (50, 50, 3, 1, 2), True)], ids=lambda nd: "%s-%s" % (nd.ns, "x".join(map(str, nd.state_shape))))
def test_parameter_layout_matches_tf_variable_order(nd):
  """the flat parameter buffer follows the reference's variable creation order and shapes"""
  lib = _lib.lib()
  spec = _spec_of(nd)
  h = C.c_void_p()
  _lib.check(lib.cpp_net_create(C.byref(spec), C.byref(h)))
  assert lib.cpp_net_num_params(h) == nd.num_params()
  shapes = nd.var_shapes()
  assert lib.cpp_net_num_vars(h) == len(shapes)
  off = 0
  for i, (_, shp) in enumerate(shapes):
    o, nd_, s4 = C.c_int64(), C.c_int32(), (C.c_int64 * 4)()
    _lib.check(lib.cpp_net_var_info(h, i, C.byref(o), C.byref(nd_), s4))
    assert o.value == off and tuple(s4[:nd_.value]) == tuple(shp)
    off += int(np.prod(shp))
  assert lib.cpp_net_feature_dim(h) == nd.feat
  assert lib.cpp_net_workspace_bytes(h, 256) > 0
  lib.cpp_net_destroy(h)


def test_survey_parameter_counts():
  """c3 actor 85,032 / critic 146,631; c5 actor 280,782 / critic 534,381"""
  assert no.ddpg_actor("a", (64, 64, 3, 1, 3), True).num_params() == 85032
  assert no.ddpg_critic("c", (64, 64, 3, 1, 3), True).num_params() == 146631
  assert no.ddpg_actor("a", (128, 128, 3, 2, 4), True).num_params() == 280782
  assert no.ddpg_critic("c", (128, 128, 3, 2, 4), True).num_params() == 534381
  assert no.naf_value("v", (64, 64, 3, 2, 3), True).num_params() == 77131


def test_invalid_specs_are_rejected():
  lib = _lib.lib()
  s = _spec_of(no.ddpg_actor("actor", (2, 2, 7), False))
  s.n_fc = 0
  h = C.c_void_p()
  assert lib.cpp_net_create(C.byref(s), C.byref(h)) == -1
  s = _spec_of(no.ddpg_actor("actor", (4, 4, 3, 1, 1), True))      # too small for three 2x2 pools
  assert lib.cpp_net_create(C.byref(s), C.byref(h)) == -1


def test_product_does_not_import_the_oracle():
  """the shipped package must never route through oracle/ (it is test infrastructure)"""
  pkg = os.path.join(ROOT, "cartpoleplusplus_b200")
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
        src = open(os.path.join(dirpath, f)).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
        assert "nets_oracle" not in src and "replay_oracle" not in src, f
