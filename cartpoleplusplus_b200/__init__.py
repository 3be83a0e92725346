"""cartpole++ RL-training hot path, B200-native (sm_100a).

Host-side mirror of the reference's Python interface (base_network / replay_memory / ddpg_cartpole /
naf_cartpole / lrpg_cartpole) over the C ABI of libcartpolepp.so (include/cartpolepp.h).
There is no CPU fallback: every numeric entry point fails loudly without the CUDA library.
"""
from ._lib import lib, CppError  # noqa: F401
