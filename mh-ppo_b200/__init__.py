"""mhppo_b200 -- B200 (sm_100a) drop-in for the hot path of BrunoudA/MH-PPO.

Host code is Python/PyTorch (device memory, streams, torch.distributed); every computation on the
path runs in hand-written CUDA kernels of `libmhppo_b200.so` behind the C ABI of `include/mhppo.h`.
There is no CPU fallback: creating an env without the built library or without a CUDA device
raises.
"""
from ._lib import lib, LibraryMissing, MhppoError, launch_count  # noqa: F401
from .vec_env import VecCrosswalkEnv, make, ENV_IDS  # noqa: F401
from .policy import Model_PPO, Adam  # noqa: F401
from .rollout import Env_rollout  # noqa: F401
from .ppo import Algo_PPO  # noqa: F401

__all__ = ["VecCrosswalkEnv", "make", "ENV_IDS", "Model_PPO", "Adam", "Env_rollout", "Algo_PPO", "lib", "LibraryMissing",
           "MhppoError", "launch_count"]
