"""gym.utils.seeding.np_random stand-in (test infrastructure only).

The reference only calls it at class-body level (Env_hybrid_multi_coop_scalable
.py:14) and never draws from the returned generator on the hot path.
"""
import numpy as np


def np_random(seed=None):
    return np.random.default_rng(seed), seed
