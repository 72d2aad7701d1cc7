"""gym.spaces stand-in (test infrastructure only): Box, Dict, Discrete."""
from collections import OrderedDict

import numpy as np


class Space:
    shape = None
    dtype = None


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = np.asarray(low)
        self.high = np.asarray(high)
        self.shape = tuple(shape) if shape is not None else self.low.shape
        self.dtype = dtype

    def sample(self):
        # the reference overwrites every key of the sampled dict immediately
        # (Env_hybrid_multi_coop_scalable.py:927-931); only the shape matters.
        return np.zeros(self.shape, dtype=self.dtype)


class Discrete(Space):
    def __init__(self, n):
        self.n = n
        self.shape = ()

    def sample(self):
        return 0


class Dict(Space):
    def __init__(self, spaces):
        if isinstance(spaces, OrderedDict):
            items = list(spaces.items())
        else:  # gym 0.26: plain dict keys are sorted
            items = sorted(spaces.items())
        self.spaces = OrderedDict(items)

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def values(self):
        return self.spaces.values()

    def items(self):
        return self.spaces.items()

    def sample(self):
        return OrderedDict((k, s.sample()) for k, s in self.spaces.items())
