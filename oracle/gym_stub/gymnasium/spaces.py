"""`gymnasium.spaces.Box` stand-in -- TEST INFRASTRUCTURE ONLY (see package docstring)."""
import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._np_random = None
        if seed is not None:
            self.seed(seed)

    @property
    def shape(self):
        return self._shape

    @property
    def np_random(self):
        if self._np_random is None:
            self.seed(None)
        return self._np_random

    def seed(self, seed=None):
        self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        return [seed]


class Box(Space):
    """Bounded box; ``sample()`` draws uniformly from the space's own generator."""

    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        shape = tuple(int(s) for s in shape)
        super().__init__(shape, dtype, seed)
        self.low = np.full(shape, low, dtype=dtype) if np.isscalar(low) else np.asarray(low, dtype=dtype).reshape(shape)
        self.high = np.full(shape, high, dtype=dtype) if np.isscalar(high) else np.asarray(high, dtype=dtype).reshape(shape)
        self.bounded_below = np.isfinite(self.low)
        self.bounded_above = np.isfinite(self.high)

    def sample(self):
        rng = self.np_random
        out = np.empty(self.shape, dtype=np.float64)
        both = self.bounded_below & self.bounded_above
        neither = ~self.bounded_below & ~self.bounded_above
        lo_only = self.bounded_below & ~self.bounded_above
        hi_only = ~self.bounded_below & self.bounded_above
        out[neither] = rng.normal(size=neither[neither].shape)
        out[lo_only] = rng.exponential(size=lo_only[lo_only].shape) + self.low[lo_only]
        out[hi_only] = -rng.exponential(size=hi_only[hi_only].shape) + self.high[hi_only]
        out[both] = rng.uniform(low=self.low[both], high=self.high[both], size=both[both].shape)
        return out.astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"
