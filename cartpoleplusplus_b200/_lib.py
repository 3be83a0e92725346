"""ctypes binding of libcartpolepp.so.  PyTorch tensors are only containers: every call passes raw
data_ptr()s and sizes (include/cartpolepp.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcartpolepp.so")
CPP_MAX_FC = 8
_lib = None


class CppError(RuntimeError):
  def __init__(self, status, msg):
    RuntimeError.__init__(self, "libcartpolepp status %d: %s" % (status, msg))
    self.status = status


class NetSpec(C.Structure):
  _fields_ = [("pixels", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
              ("input_dim", C.c_int32), ("n_fc", C.c_int32),
              ("fc_out", C.c_int32 * CPP_MAX_FC), ("fc_act", C.c_int32 * CPP_MAX_FC),
              ("concat_at", C.c_int32), ("action_dim", C.c_int32), ("fc_dropout", C.c_int32 * CPP_MAX_FC), ("batch_norm", C.c_int32)]


class DDPGConfig(C.Structure):
  _fields_ = [("actor", NetSpec), ("critic", NetSpec),
              ("actor_lr", C.c_float), ("critic_lr", C.c_float), ("discount", C.c_float),
              ("gradient_clip", C.c_float), ("target_update_rate", C.c_float),
              ("max_batch", C.c_int32), ("world_size", C.c_int32), ("rank", C.c_int32)]


class DDPGBuffers(C.Structure):
  _fields_ = [("params", C.c_void_p), ("target_params", C.c_void_p), ("grads", C.c_void_p),
              ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64)]


class NAFConfig(C.Structure):
  _fields_ = [("value", NetSpec), ("mu", NetSpec), ("l", NetSpec),
              ("discount", C.c_float), ("gradient_clip", C.c_float), ("target_update_rate", C.c_float),
              ("optimiser", C.c_int32),
              ("lr", C.c_float), ("momentum", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
              ("max_batch", C.c_int32), ("action_dim", C.c_int32), ("world_size", C.c_int32), ("rank", C.c_int32),
              ("share_input_state_representation", C.c_int32)]


class NAFBuffers(C.Structure):
  _fields_ = [("params", C.c_void_p), ("target_params", C.c_void_p), ("grads", C.c_void_p),
              ("slots", C.c_void_p), ("opt_state", C.c_void_p),
              ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64)]


class LRPGConfig(C.Structure):
  _fields_ = [("model", NetSpec), ("gradient_clip", C.c_float), ("optimiser", C.c_int32),
              ("lr", C.c_float), ("momentum", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
              ("max_batch", C.c_int32)]


class LRPGBuffers(C.Structure):
  _fields_ = [("params", C.c_void_p), ("grads", C.c_void_p), ("slots", C.c_void_p), ("opt_state", C.c_void_p),
              ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64)]


# every symbol include/cartpolepp.h declares (tests/test_abi.py checks the list against the header)
SYMBOLS = """cpp_version cpp_last_error cpp_launch_count cpp_piece_overflow_count cpp_set_option
cpp_conv_wgrad_mma_scratch_bytes cpp_conv_wgrad_mma cpp_conv_dgrad_tc_scratch_bytes cpp_conv_dgrad_tc cpp_ddpg_step_backward cpp_ddpg_step_apply cpp_ddpg_train_step
cpp_conv_forward cpp_conv_dgrad cpp_conv_wgrad_scratch_floats cpp_conv_wgrad
cpp_conv_tc_scratch_bytes cpp_conv_forward_tc
cpp_mt_create cpp_mt_destroy cpp_mt_seed cpp_mt_set_state cpp_mt_get_state cpp_mt_randint
cpp_replay_gather cpp_slot_stats cpp_moments_from_slots cpp_moments_scratch_doubles cpp_channel_moments
cpp_net_create cpp_net_destroy cpp_net_num_params cpp_net_num_vars cpp_net_var_info cpp_net_feature_dim
cpp_net_workspace_bytes cpp_net_forward cpp_net_backward
cpp_norm_scratch_doubles cpp_global_norm_scale cpp_optimiser_apply cpp_soft_update
cpp_ddpg_create cpp_ddpg_destroy cpp_ddpg_workspace_bytes cpp_ddpg_layout cpp_ddpg_bind cpp_ddpg_set_moments
cpp_ddpg_actor_backward cpp_ddpg_actor_apply cpp_ddpg_actor_train cpp_ddpg_critic_backward cpp_ddpg_critic_apply
cpp_ddpg_critic_train cpp_ddpg_check_loss cpp_ddpg_action_given cpp_ddpg_update_targets cpp_ddpg_debug_view cpp_nccl_unique_id cpp_nccl_version cpp_ddpg_comm_init cpp_naf_comm_init cpp_ddpg_p2p_prepare cpp_ddpg_p2p_connect cpp_naf_p2p_prepare cpp_naf_p2p_connect cpp_ddpg_all_reduce_grads cpp_naf_all_reduce_grads cpp_ddpg_action_given_fast cpp_naf_action_given_fast
cpp_naf_create cpp_naf_destroy cpp_naf_workspace_bytes cpp_naf_layout cpp_naf_bind cpp_naf_set_moments
cpp_naf_backward cpp_naf_apply cpp_naf_train cpp_naf_debug_values cpp_naf_action_given cpp_naf_value_given
cpp_naf_update_targets cpp_naf_debug_view
cpp_lrpg_create cpp_lrpg_destroy cpp_lrpg_workspace_bytes cpp_lrpg_num_params cpp_lrpg_bind cpp_lrpg_train
cpp_lrpg_logits""".split()

_INT64_RET = {"cpp_launch_count", "cpp_piece_overflow_count", "cpp_conv_dgrad_tc_scratch_bytes", "cpp_conv_wgrad_mma_scratch_bytes", "cpp_conv_tc_scratch_bytes", "cpp_conv_wgrad_scratch_floats", "cpp_moments_scratch_doubles", "cpp_net_num_params", "cpp_net_workspace_bytes", "cpp_norm_scratch_doubles",
              "cpp_ddpg_workspace_bytes", "cpp_naf_workspace_bytes", "cpp_lrpg_workspace_bytes", "cpp_lrpg_num_params"}


def lib():
  """Load the CUDA library; there is NO fallback - a missing library is a hard error."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise ImportError("%s is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                      "This package has no CPU fallback." % LIB_PATH)
  l = C.CDLL(LIB_PATH)
  for name in SYMBOLS:
    fn = getattr(l, name)           # AttributeError if the library does not export it
    fn.restype = C.c_int64 if name in _INT64_RET else C.c_int
  l.cpp_last_error.restype = C.c_char_p
  _lib = l
  return l


def check(status):
  if status != 0:
    raise CppError(status, lib().cpp_last_error().decode("utf-8", "replace"))
  return status


def ptr(t):
  """device (or host) pointer of a torch tensor / numpy array as c_void_p; None -> NULL"""
  if t is None:
    return C.c_void_p(0)
  if hasattr(t, "data_ptr"):
    return C.c_void_p(t.data_ptr())
  return C.c_void_p(t.ctypes.data)


def ptr_array(tensors):
  """host array of device pointers (const T* const*)"""
  return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def stream_ptr():
  import torch
  return C.c_void_p(torch.cuda.current_stream().cuda_stream)
