import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.dtype = low, high, dtype
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self._rng = np.random.default_rng()

    def seed(self, s=None):
        self._rng = np.random.default_rng(s)

    def sample(self):
        return self._rng.uniform(-1, 1, size=self.shape).astype(self.dtype)
