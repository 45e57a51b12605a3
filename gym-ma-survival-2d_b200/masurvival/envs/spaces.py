"""Minimal stand-ins for the `gym.spaces` classes the reference's env exposes
(masurvival/envs/masurvival_env.py:391-453): just enough for callers that
read `.shape`, index a Dict, call `.contains()` or `.sample()`."""
import numpy as np


class Space:
    shape = None

    def contains(self, x):
        raise NotImplementedError

    def sample(self):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)


class Box(Space):
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def sample(self):
        return np.random.standard_normal(self.shape).astype(self.dtype)


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.shape = self.nvec.shape

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

    def sample(self):
        return np.random.randint(0, self.nvec).astype(np.int64)


class Tuple(Space):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)

    def contains(self, x):
        return len(x) == len(self.spaces) and all(s.contains(v) for s, v in zip(self.spaces, x))

    def sample(self):
        return tuple(s.sample() for s in self.spaces)

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)


class Dict(Space):
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def contains(self, x):
        return set(x.keys()) == set(self.spaces.keys()) and all(self.spaces[k].contains(v) for k, v in x.items())

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()
