"""Minimal stand-in for the `gym` 0.26 package (TEST INFRASTRUCTURE ONLY).

The reference envs (`/root/reference/Environments/*.py`) import gym for exactly
four things: the `Env` base class, `spaces.{Box,Dict,Discrete}`, `utils.seeding
.np_random` and `envs.registration.register` / `gym.make`.  gym is not installed
in this image, so the oracle harness puts this directory on `sys.path` to import
the UNMODIFIED reference.  Nothing under `mh-ppo_b200/` imports it.

Behaviour that matters and is reproduced: `spaces.Dict` sorts plain-dict keys
(gym 0.26), which fixes the flat observation order `car, (car_follow), env, ped`
that the reference rollout indexes (Coop-MH-PPO-scalable.py:549-554).
"""
from . import spaces, utils, envs  # noqa: F401
from .envs.registration import register, make  # noqa: F401


class Env:
    metadata = {}

    def __init__(self, *a, **k):
        pass

    def reset(self, seed=None, options=None):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    @property
    def unwrapped(self):
        return self
