"""Space stand-ins: only construction, .contains/.sample basics and attribute access."""
from __future__ import annotations

import numpy as np

from .utils import seeding


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else (dtype if dtype is str else np.dtype(dtype))
        self._np_random = None
        if seed is not None:
            if isinstance(seed, np.random.Generator):
                self._np_random = seed
            else:
                self.seed(seed)

    def __class_getitem__(cls, item):
        return cls

    @property
    def shape(self):
        return self._shape

    @property
    def np_random(self):
        if self._np_random is None:
            self.seed()
        return self._np_random

    def seed(self, seed=None):
        self._np_random, s = seeding.np_random(seed)
        return s

    def sample(self, mask=None):
        raise NotImplementedError

    def contains(self, x):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        self.n = int(n)
        self.start = int(start)
        super().__init__((), np.int64, seed)

    def sample(self, mask=None):
        return int(self.start + self.np_random.integers(self.n))

    def contains(self, x):
        try:
            xi = int(x)
        except Exception:
            return False
        return self.start <= xi < self.start + self.n

    def __eq__(self, other):
        return isinstance(other, Discrete) and (self.n, self.start) == (other.n, other.start)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        dtype = np.dtype(dtype)
        if shape is None:
            shape = np.shape(low)
        self.low = np.full(shape, low, dtype=dtype) if np.isscalar(low) else np.asarray(low, dtype=dtype)
        self.high = np.full(shape, high, dtype=dtype) if np.isscalar(high) else np.asarray(high, dtype=dtype)
        super().__init__(shape, dtype, seed)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def sample(self, mask=None):
        return self.np_random.integers(self.low, self.high, endpoint=True).astype(self.dtype)


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None):
        self.nvec = np.asarray(nvec, dtype=dtype)
        super().__init__(self.nvec.shape, dtype, seed)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))


class Text(Space):
    def __init__(self, max_length, min_length=1, charset=None, seed=None):
        self.max_length = max_length
        self.min_length = min_length
        super().__init__(dtype=str, seed=seed)

    def contains(self, x):
        return isinstance(x, str)


class Dict(Space):
    def __init__(self, spaces=None, seed=None, **kw):
        self.spaces = dict(spaces or {})
        self.spaces.update(kw)
        super().__init__(None, None, seed)

    def __getitem__(self, k):
        return self.spaces[k]

    def __setitem__(self, k, v):
        self.spaces[k] = v

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()

    def contains(self, x):
        return isinstance(x, dict) and all(k in x for k in self.spaces)
