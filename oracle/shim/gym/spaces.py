"""gym.spaces shim (see gym/__init__.py)."""
import numpy as np

_rng = np.random.default_rng(0)


class Space:
    shape = None

    def contains(self, x):
        raise NotImplementedError

    def sample(self):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

    def contains(self, x):
        x = np.asarray(x)
        return (np.can_cast(x.dtype, self.dtype) and x.shape == self.shape
                and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high)))

    def sample(self):
        return _rng.standard_normal(self.shape).astype(self.dtype)


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.shape = self.nvec.shape

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= 0)) and bool(np.all(x < self.nvec))

    def sample(self):
        return (_rng.random(self.nvec.shape) * self.nvec).astype(np.int64)


class Tuple(Space):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)

    def contains(self, x):
        return isinstance(x, (tuple, list)) and len(x) == len(self.spaces) and \
            all(s.contains(v) for s, v in zip(self.spaces, x))

    def sample(self):
        return tuple(s.sample() for s in self.spaces)

    def __getitem__(self, i):
        return self.spaces[i]

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)


class Dict(Space):
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def contains(self, x):
        return isinstance(x, dict) and len(x) == len(self.spaces) and \
            all(k in self.spaces and self.spaces[k].contains(v) for k, v in x.items())

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()
