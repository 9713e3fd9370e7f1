import numpy as np


class _Space:
    def __init__(self):
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)


class Discrete(_Space):
    def __init__(self, n):
        super().__init__()
        self.n = int(n)

    def sample(self):
        return int(self._rng.integers(0, self.n))


class MultiDiscrete(_Space):
    def __init__(self, nvec):
        super().__init__()
        self.nvec = np.asarray(nvec, dtype=np.int64)

    def sample(self):
        return (self._rng.random(self.nvec.shape) * self.nvec).astype(np.int64)


class Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        super().__init__()
        self.low, self.high, self.dtype = low, high, dtype
        self.shape = tuple(shape) if shape is not None else np.shape(low)


class Dict(_Space):
    def __init__(self, spaces=None, **kw):
        super().__init__()
        self.spaces = dict(spaces or {}, **kw)
